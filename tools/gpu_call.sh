mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_match.py -m gpu -q -x -k "match_all" > gpurun_out/r02ak_memcheck.txt 2>&1
echo "rc=$?" >> gpurun_out/r02ak_memcheck.txt
tail -n 12 gpurun_out/r02ak_memcheck.txt
