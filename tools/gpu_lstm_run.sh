python -m pytest tests/test_gpu_encoder.py -x -q -k "lstm or encoder or compute" > gpurun_out/t8.log 2>&1
XB_LSTM_VARIANT=1 python -m pytest tests/test_gpu_encoder.py -x -q -k "lstm or encoder or compute" > gpurun_out/t8b.log 2>&1
python tools/lstm_timeline.py 512 > gpurun_out/timeline4.log 2>&1
XB_LSTM_VARIANT=1 python tools/lstm_timeline.py 512 >> gpurun_out/timeline4.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
XB_LSTM_VARIANT=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5b.json 2> gpurun_out/bench5b.err
