python tools/i8_check.py persist > gpurun_out/i8_check10.txt 2>&1; echo rc=$?
grep "persistent ==" gpurun_out/i8_check10.txt
for LA in 1 0; do for W in 128 64 32 16 8; do echo -n "lookahead=$LA W=$W: "; AGPN_LOOKAHEAD=$LA python tools/sweep.py 2048 $W | cut -c40-100; done; done
for LA in 1 0; do echo -n "lookahead=$LA C4 W=4: "; AGPN_LOOKAHEAD=$LA SWEEP_Q=3 python tools/sweep.py 8192 4 2 | cut -c40-100; done
