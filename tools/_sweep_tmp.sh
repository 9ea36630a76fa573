for C in 24 37 50 74; do for S in 4 6 8; do echo -n "clusters=$C streams=$S: "; AGPN_I8_CLUSTERS=$C AGPN_STREAMS=$S python tools/sweep.py 2048 16 | cut -c40-100; done; done
for G in 16 8 4; do echo -n "group=$G: "; AGPN_GROUP=$G python tools/sweep.py 2048 16 | cut -c40-100; done
