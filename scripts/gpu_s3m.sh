SMCB200_DBG_STEP=64 timeout 100 python scripts/quick_time.py 10,16 500 2>&1 | grep -E "k_filter_small|best" | awk 'NR%6==1 || /best/' | cut -c1-250
