mkdir -p gpurun_out
L=gpurun_out/s3x.log
./tests/cpp/user_csmc 1000 40 7 2>&1 | cut -c1-160 | tail -8 > $L
timeout 900 python -m pytest tests/test_user_csmc.py tests/test_gpu_csmc.py tests/test_user_model.py -m gpu -x -q 2>&1 | tail -12 >> $L
cat $L
