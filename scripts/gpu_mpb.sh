for M in 1 2 4; do echo "MPB=$M"; ERTDIFF_CHAIN_MPB=$M python scripts/chain_sweep.py --members 148,256,296,512,1024 --precisions fp32 2>&1 | grep chain_ms; done
