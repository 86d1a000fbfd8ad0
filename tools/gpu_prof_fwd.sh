B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean"
for g in 592 1184; do
  EPI_FWD_SEGMENTS=7 EPI_FWD_GRID=$g timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_forward_seg -s 2 -c 1 -o gpurun_out/prof_fwd_g$g -f $B > gpurun_out/ncu_fwd_g$g.log 2>&1
  tail -1 gpurun_out/ncu_fwd_g$g.log
done
