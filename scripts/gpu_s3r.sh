for d in 0 128 256; do SMCB200_DBG_STEP=$d SMCB200_FORCE_ISLAND=1 timeout 100 python scripts/ab_time.py 24 100 | cut -c1-150; done
