# int8 digit-plane mode of the tcgen05 path: parity first, then throughput on C3 / C4
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_i8.py -m gpu -x -q -s ) > gpurun_out/r2b_i8_test.log 2>&1
echo "i8 tests rc=$?" 
tail -5 gpurun_out/r2b_i8_test.log
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k c3_depth ) > gpurun_out/r2b_c3depth_test.log 2>&1
echo "c3 depth rc=$?"
tail -8 gpurun_out/r2b_c3depth_test.log
for w in c3 c4; do for prec in i8x3 i8x2; do
  timeout 300 python bench.py --workload $w --prec $prec --no-cpu-baseline > gpurun_out/r2b_bench_${w}_${prec}.json 2> gpurun_out/r2b_bench_${w}_${prec}.err
  echo "$w $prec rc=$?"; head -c 600 gpurun_out/r2b_bench_${w}_${prec}.json; echo
done; done
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowshard.py -m gpu -x -q ) > gpurun_out/r2b_parity_test.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r2b_parity_test.log
( timeout 600 python -m pytest tests/test_gpu_threads.py -m gpu -x -q ) > gpurun_out/r2b_threads_test.log 2>&1
echo "threads rc=$?"; tail -5 gpurun_out/r2b_threads_test.log
