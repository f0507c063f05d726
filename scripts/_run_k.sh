timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_integrate.py -m gpu -x -q -k "batch_equals_sequential or prepared_jobs or freespace_and_edge or far_points" 2>&1 | tail -15
echo "memcheck rc=$?"
