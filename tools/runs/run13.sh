mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputest23.log 2>&1; tail -15 gpurun_out/gputest23.log | cut -c1-300
