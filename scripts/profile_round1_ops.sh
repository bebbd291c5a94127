# ncu --set full of the secondary kernels (one launch each, after warm-up)
set -x
prof() {  # name, kernel regex, script arg, skip count
  python scripts/one_ops.py $3 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -o gpurun_out/prof_$1 python scripts/one_ops.py $3 > gpurun_out/ncu_$1.log 2>&1
  tail -n 1 gpurun_out/ncu_$1.log
}
prof bq ball_query_scan bq 2
prof fps fps_d3 fps 2
prof chamfer_fwd chamfer_fwd chamfer 4
prof chamfer_bwd chamfer_bwd chamfer 4
prof knn_k1 knn_prune chamfer 4
prof knn_backward knn_backward knnbwd 2
prof gather 'gather_kernel' knnbwd 2
TC_N=4 python scripts/one_tc.py > gpurun_out/plain_rr.log 2>&1 && TC_N=4 ncu --set full --clock-control none --import-source on -k regex:knn_tc_rerank -s 1 -c 1 -o gpurun_out/prof_tc_rerank python scripts/one_tc.py > gpurun_out/ncu_rr.log 2>&1
tail -n 1 gpurun_out/ncu_rr.log
