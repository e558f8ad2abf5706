timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 300 -x -k "tc or bip" 2>&1 | tail -3
for prec in bf16x1 bf16x3; do for f in 0 1 2 99; do echo "prec=$prec fine=$f"; ISB_TC_FINE=$f timeout 300 python bench.py --workload c4 --prec $prec --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"; done; done
