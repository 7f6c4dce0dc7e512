#!/bin/bash
# box probe + compute-sanitizer passes; outputs under gpurun_out/
mkdir -p gpurun_out
{ free -g; nproc; lscpu | grep -E "Model name|Thread|Core|Socket|NUMA"; cat /sys/fs/cgroup/memory.max 2>/dev/null; df -h /dev/shm /tmp; nvidia-smi -L; which mpirun mpicxx; python -c "import numpy, glob, os; print(glob.glob(os.path.join(os.path.dirname(numpy.__file__), '..', 'numpy.libs', '*')))"; } > gpurun_out/r02_box.txt 2>&1
K="test_ttm_first or test_ttm_multi or test_sym_eig_topk_large"
timeout 700 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python -m pytest tests/test_kernels_gpu.py -x -q -k "$K" > gpurun_out/r02_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_box.txt
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck.log python -m pytest tests/test_kernels_gpu.py -x -q -k "$K" > gpurun_out/r02_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_box.txt
tail -3 gpurun_out/r02_memcheck_pytest.log gpurun_out/r02_racecheck_pytest.log
tail -5 gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log
cat gpurun_out/r02_box.txt
