#!/bin/bash
# compute-sanitizer over the GPU test suite at small shapes.  Run under gpurun; logs in gpurun_out/r2.
O=gpurun_out/r2
mkdir -p $O
T="tests/test_gpu_kernels.py tests/test_gpu_knn.py tests/test_gpu_wnn.py tests/test_gpu_mofa.py tests/test_gpu_tfidf.py"
K="not config0 and not scale"
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/sanitizer_memcheck.log python -m pytest $T tests/test_gpu_round2.py -q -x -k "$K and not slice and not stager and not to_device") > $O/sanitizer_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> $O/sanitizer_memcheck_pytest.log
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/sanitizer_racecheck.log python -m pytest tests/test_gpu_kernels.py tests/test_gpu_knn.py tests/test_gpu_wnn.py tests/test_gpu_round2.py -q -x -k "$K and not slice and not stager and not to_device and not schedules and not host_path") > $O/sanitizer_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> $O/sanitizer_racecheck_pytest.log
tail -3 $O/sanitizer_memcheck_pytest.log $O/sanitizer_racecheck_pytest.log; tail -5 $O/sanitizer_memcheck.log $O/sanitizer_racecheck.log
