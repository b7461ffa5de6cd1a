# ncu evidence for the current kernels: launch list of one configs[1] step + full captures of the top kernels
python tools/profile_step.py 512 2 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 512 2 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_persistent -s 5 -c 1 -o gpurun_out/prof_lstm python tools/profile_step.py 512 2 > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inproj_kernel -s 5 -c 2 -o gpurun_out/prof_inproj python tools/profile_step.py 512 2 > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/prof_gemm python tools/profile_step.py 512 2 > gpurun_out/ncu5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:crf_ -s 3 -c 3 -o gpurun_out/prof_crf python tools/profile_step.py 512 2 > gpurun_out/ncu2.log 2>&1
echo finished > gpurun_out/profile_done.log
