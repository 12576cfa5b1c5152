#!/bin/bash
# evidence run (one GPU, one gpurun call): tests, bench lines, ncu launch list + full capture + source pages, sanitizer; results under gpurun_out/evidence, summarised into profiles/ with tools/ncu_report.py and tools/sass_excerpt.py
cd $GRAFT_REPO_ROOT
O=gpurun_out/evidence; mkdir -p $O; T=/tmp/prof2; mkdir -p $T
timeout 500 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt; tail -n 3 $O/pytest_gpu.txt
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --stress --no-extras --no-cpu-baseline > $O/bench_stress.json 2>/dev/null; echo "stress rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>/dev/null; echo "reference rc=$?"
K='regex:(ranges_seed|ranges|hist|remap|moments|apply)_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file $O/launches_idt.csv python bench.py --steps 1 --warmup 3 --passes 2 --frames 4 --kernels-only > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file $O/launches_linear.csv python bench.py --linear-only --linear-pairs 64 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 66 -c 11 -o $T/prof_idt python bench.py --steps 1 --warmup 3 --passes 2 --frames 4 --kernels-only > $O/ncu_idt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 8 -c 8 -o $T/prof_linear python bench.py --linear-only --linear-pairs 64 > $O/ncu_linear.log 2>&1
ncu -i $T/prof_idt.ncu-rep --page raw --csv > $O/raw_idt.csv 2>/dev/null
ncu -i $T/prof_linear.ncu-rep --page raw --csv > $O/raw_linear.csv 2>/dev/null
ncu -i $T/prof_idt.ncu-rep --page source --csv -k regex:hist_kernel --launch-skip 1 --launch-count 1 > $O/src_hist1.csv 2>/dev/null
ncu -i $T/prof_idt.ncu-rep --page source --csv -k regex:remap_kernel --launch-skip 1 --launch-count 1 > $O/src_remap1.csv 2>/dev/null
ncu -i $T/prof_idt.ncu-rep --page source --csv -k regex:"ranges_kernel" --launch-skip 1 --launch-count 1 > $O/src_ranges4.csv 2>/dev/null
ncu -i $T/prof_linear.ncu-rep --page source --csv -k regex:apply_kernel --launch-skip 0 --launch-count 1 > $O/src_apply_lab.csv 2>/dev/null
ncu -i $T/prof_linear.ncu-rep --page source --csv -k regex:moments_kernel --launch-skip 0 --launch-count 1 > $O/src_moments_lab.csv 2>/dev/null
gzip -f $O/src_*.csv
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_workload.py > $O/san_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -n 2 $O/san_memcheck.txt
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitizer_workload.py > $O/san_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -n 2 $O/san_racecheck.txt
du -sh $O
