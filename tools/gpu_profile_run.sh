python -m pytest tests -m gpu -x -q > gpurun_out/t7.log 2>&1
python tools/profile_step.py 512 2 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 512 2 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:crf_backward -s 1 -c 1 -o gpurun_out/prof_crf_backward python tools/profile_step.py 512 2 > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_persistent -s 5 -c 1 -o gpurun_out/prof_lstm python tools/profile_step.py 512 2 > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 12 -c 2 -o gpurun_out/prof_gemm python tools/profile_step.py 512 2 > gpurun_out/ncu4.log 2>&1
echo finished >> gpurun_out/t7.log
