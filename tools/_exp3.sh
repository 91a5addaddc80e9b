python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
for v in 1 0; do
echo "== VQA_TAIL_SPLIT=$v"
VQA_TAIL_SPLIT=$v python tools/per_op_ms.py 256 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for o in d['ops']:
    if 'tail' in o['name']: print(o['name'], o['kernel'], round(o['ms']*1e3,1))
"
VQA_TAIL_SPLIT=$v python bench.py --quick --steps 30 2>&1 | tail -1
done
