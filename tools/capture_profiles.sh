#!/bin/bash
# Round-end measurement run on the GPU box: tests, bench lines, ncu launch list and `ncu --set full` summaries (text only: the
# .ncu-rep files are summarised on the box with tools/ncu_summary.py / tools/ncu_lines.py and deleted, gpurun_out/ is capped at 64 MiB).
# Usage (through gpurun): bash tools/capture_profiles.sh r02
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $out/${tag}_tests.log
python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gapped-stress --no-workloads --no-whole-reads"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_bench.csv $BENCH1 > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:scan_ -c 2 -f -o /tmp/scan $BENCH1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/scan.ncu-rep 625000 0 > $out/${tag}_ncu_scan_c2_10Mpairs.txt 2>&1
python tools/ncu_lines.py /tmp/scan.ncu-rep scan_fx 625000 0.012 > $out/${tag}_ncu_scan_fx_lines.txt 2>&1
python tools/ncu_lines.py /tmp/scan.ncu-rep "scan_kernel" 67300 0.015 > $out/${tag}_ncu_scan_general_lines.txt 2>&1
ncu --set full --clock-control none -k regex:sw_pairs -c 2 -f -o /tmp/swp5 python tools/run_workload.py --workload c5 --pairs 200000 --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/swp5.ncu-rep > $out/${tag}_ncu_sw_pairs_c5_200kpairs.txt 2>&1
ncu --set full --clock-control none -k regex:sw_all -c 1 -f -o /tmp/swa python tools/run_workload.py --workload c3b --pairs 100000 --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/swa.ncu-rep > $out/${tag}_ncu_sw_all_c3b_100kpairs.txt 2>&1
ncu --set full --clock-control none -k regex:sw_pairs -c 1 -f -o /tmp/swp3 python tools/run_workload.py --workload c3 --pairs 4000000 --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/swp3.ncu-rep > $out/${tag}_ncu_sw_pairs_c3_4Mpairs.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:sw_pairs -c 1 -f -o /tmp/swp4 python tools/run_workload.py --workload c4 --pairs 4000000 --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/swp4.ncu-rep > $out/${tag}_ncu_sw_pairs_packed_c4_4Mpairs.txt 2>&1
python tools/ncu_lines.py /tmp/swp4.ncu-rep sw_pairs 1 0.01 > $out/${tag}_ncu_sw_pairs_packed_c4_lines.txt 2>&1
ncu --set full --clock-control none -k regex:sw_all -c 1 -f -o /tmp/swa5 python tools/run_workload.py --workload c5 --pairs 200000 --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/swa5.ncu-rep > $out/${tag}_ncu_sw_all_c5_200kpairs.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:cand5_thread -c 1 -f -o /tmp/c5t python tools/run_workload.py --workload c4 --pairs 4000000 --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/c5t.ncu-rep > $out/${tag}_ncu_cand5_thread_c4_4Mpairs.txt 2>&1
python tools/ncu_lines.py /tmp/c5t.ncu-rep cand5_thread 1 0.02 > $out/${tag}_ncu_cand5_thread_lines.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_c4.csv python tools/run_workload.py --workload c4 --pairs 10000000 --iters 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_c5.csv python tools/run_workload.py --workload c5 --pairs 1000000 --iters 2 > /dev/null 2>&1
tools/sw_wavefront_baseline > $out/${tag}_sw_wavefront_baseline.json 2>&1
du -sh $out
