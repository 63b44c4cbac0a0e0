cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
bash tools/profile_gpu.sh 65536
python tools/bench_configs.py > gpurun_out/r2m_configs.log 2>&1; echo "configs rc=$?"
tail -5 gpurun_out/r2m_pytest_gpu.log
cut -c1-600 gpurun_out/r2m_bench.json
