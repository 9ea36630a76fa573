for N in 256 384 512 768 1024 1536; do for M in i8 dmma; do echo -n "N=$N mode=$M: "; AGPN_SYRK=$M python tools/sweep.py $N 256 | cut -c30-110; done; done
