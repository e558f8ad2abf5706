timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sparse.py -m gpu -q --timeout 300 -x -k "tc or bip or per_replica" 2>&1 | tail -2
for w in c4 c3; do for prec in bf16x1 bf16x3; do echo "$w $prec"; timeout 300 python bench.py --workload $w --prec $prec --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"; done; done
