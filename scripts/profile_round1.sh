set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python scripts/one_knn.py 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_prune -s 1 -c 1 -o gpurun_out/prof_knn_prune_r1c python scripts/one_knn.py 16 > gpurun_out/ncu7.log 2>&1
TC_N=4 python scripts/one_tc.py > gpurun_out/plain2.log 2>&1 && \
TC_N=4 ncu --set full --clock-control none --import-source on -k regex:knn_tc_scan -s 1 -c 1 -o gpurun_out/prof_knn_tc_r1c python scripts/one_tc.py > gpurun_out/ncu8.log 2>&1
for f in gpurun_out/ncu7.log gpurun_out/ncu8.log gpurun_out/ncu_l.log; do tail -n 2 $f; done
