mkdir -p gpurun_out/r2
timeout 900 python -m pytest -m gpu tests/test_stream_gpu.py tests/test_synthesis_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/r2/range_test.log 2>&1; tail -5 gpurun_out/r2/range_test.log
