for W in 16 128; do for G in 8 16; do AGPN_SYRK=i8 AGPN_GROUP=$G python tools/sweep.py 2048 $W; done; done
AGPN_GROUP=8 python tools/sweep.py 2048 16
for G in 8 16 64; do AGPN_SYRK=i8 AGPN_GROUP=$G SWEEP_Q=3 python tools/sweep.py 8192 4 2; done
AGPN_GROUP=8 SWEEP_Q=3 python tools/sweep.py 8192 4 2
for G in 8 64; do AGPN_SYRK=i8 AGPN_GROUP=$G SWEEP_Q=3 python tools/sweep.py 8192 32 1; done
AGPN_GROUP=8 SWEEP_Q=3 python tools/sweep.py 8192 32 1
