for pf in 0 1 0 1; do
CG_MERGE_BULK_PREFETCH=$pf python bench.py --config C5 --c3-submaps 16 --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/c5_pf$pf.json 2> gpurun_out/c5_pf.err || tail -3 gpurun_out/c5_pf.err
python -c "
import json;d=json.loads(open('gpurun_out/c5_pf$pf.json').read().strip().splitlines()[-1]);print('prefetch=$pf', d['ms_per_step'], d['value'], d['stages_ms_per_step'])"
done
python -m pytest tests/test_gpu_merge.py tests/test_gpu_reproject.py -m gpu -x -q 2>&1 | tail -3
CG_MERGE_BULK_PREFETCH=1 python -m pytest tests/test_gpu_merge.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
