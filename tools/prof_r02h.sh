CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-metrics --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r02h_plain.json 2> gpurun_out/r02h_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02h_launches.csv $CMD > gpurun_out/r02h_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_kd_subtree|k_transfer_fwd|k_transfer_bwd_search|k_accumulate|k_reproject|k_filter|k_to_rgb8" -s 12 -c 12 -o gpurun_out/r02h_full_a -f $CMD > gpurun_out/r02h_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_g_count|k_g_stage|k_g_apply1|k_g_pass2" -s 80 -c 8 -o gpurun_out/r02h_full_b -f $CMD > gpurun_out/r02h_ncu3.log 2>&1
grep -c . gpurun_out/r02h_launches.csv; grep "Profiling" gpurun_out/r02h_ncu2.log gpurun_out/r02h_ncu3.log | cut -c1-80
