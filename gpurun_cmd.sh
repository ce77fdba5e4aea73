mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "recycle or lossless" 2>&1 | tail -3
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or step_host or graph_replayable or recycle" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
