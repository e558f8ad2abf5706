for l in isingmodel.jl_b200/libising_b200.so scratch_ab/libk1.so; do for prec in bf16x1 bf16x3; do echo "$l $prec"; ISING_B200_LIB=$PWD/$l timeout 300 python bench.py --workload c4 --prec $prec --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"; done; done
