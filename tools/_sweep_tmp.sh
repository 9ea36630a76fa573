for L in "" artgpn_b200/variants/lib_pf4.so artgpn_b200/variants/lib_pf8.so; do
  echo "lib=$L"
  AGPN_LIB=$L python tools/sweep.py 2048 128
  AGPN_LIB=$L python tools/sweep.py 2048 16
done
for S in 1 2 3 4 8; do AGPN_STREAMS=$S python tools/sweep.py 2048 128; AGPN_STREAMS=$S python tools/sweep.py 2048 16; done
