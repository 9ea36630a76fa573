for F in 1 0; do for S in 6 8; do echo "fuse=$F streams=$S"; AGPN_I8_FUSE=$F AGPN_STREAMS=$S python tools/sweep.py 2048 128; AGPN_I8_FUSE=$F AGPN_STREAMS=$S python tools/sweep.py 2048 16; done; done
for C in 60 66 70; do echo "clusters=$C"; AGPN_I8_CLUSTERS=$C python tools/sweep.py 2048 128; done
