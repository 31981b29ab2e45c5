# ncu --set full summaries of the widened rows' kernels (round 2)
mkdir -p gpurun_out
run() { # name regex skip count
  timeout 120 python scripts/prof_rows.py $1 > gpurun_out/r2i_$1_plain.log 2>&1 || { echo "plain run of $1 failed"; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r2i_$1 python scripts/prof_rows.py $1 > gpurun_out/r2i_$1_ncu.log 2>&1
  tail -n 1 gpurun_out/r2i_$1_ncu.log
  python scripts/ncu_summary.py gpurun_out/r2i_$1.ncu-rep gpurun_out/r2i_$1 > /dev/null 2>&1
  rm -f gpurun_out/r2i_$1.ncu-rep
}
run lineart "k_step" 10 2
run blockpf "k_bpf_step|k_bpf_collect" 10 3
run linreg "k_la_mcmc|k_la_cess|k_la_cov" 6 4
run csmc "k_csmc_chain" 0 1
run small "k_filter_small" 0 1
run draws "k_multinomial_draw|k_cum_tiles|k_cum_write|k_cnt_tiles|k_cnt_scan|k_draw_buckets" 6 6
