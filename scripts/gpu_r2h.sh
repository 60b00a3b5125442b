#!/bin/bash
# Round 2, call H: the whole GPU suite at HEAD, the contract bench line + the other BASELINE configs at N=1, smoke.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,temperature.gpu --format=csv > $O/r2h_smi.txt 2>&1
timeout 2700 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/r2h_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" $O/r2h_pytest.log | tail -n 2; grep -E "^FAILED|^E  " $O/r2h_pytest.log | cut -c1-200 | head -20
timeout 900 python bench.py --ops-json $O/r2h_ops.json > $O/r2h_bench.log 2> $O/r2h_bench.err
echo "bench exit $?"; tail -n 1 $O/r2h_bench.log | cut -c1-400
for c in 3 4 5 1; do
  timeout 900 python bench.py --config $c --no-cpu-baseline > $O/r2h_bench_c$c.log 2> $O/r2h_bench_c$c.err
  echo "bench config $c exit $?"; tail -n 1 $O/r2h_bench_c$c.log | cut -c1-300
done
timeout 900 python bench.py --precision high --batch 2 --no-cpu-baseline > $O/r2h_bench_high.log 2> $O/r2h_bench_high.err
echo "bench high exit $?"; tail -n 1 $O/r2h_bench_high.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
