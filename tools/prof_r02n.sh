# final-state captures of the ingest kernels and of the metrics kernels (round 2, r02n)
python tools/ingest_prof.py > gpurun_out/r02n_ingest_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_chroma_420_to_444|k_luma_to_16|k_widen_u8|k_geometry_set|k_occupancy_convert" -s 12 -c 5 -o gpurun_out/r02n_ingest -f python tools/ingest_prof.py > gpurun_out/r02n_ncu1.log 2>&1
python tools/met_prof.py > gpurun_out/r02n_met_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_nn_near|k_nn_pending|k_nn_far|k_column_|k_import_clouds" -s 330 -c 16 -o gpurun_out/r02n_met -f python tools/met_prof.py > gpurun_out/r02n_ncu2.log 2>&1
grep "Profiling" gpurun_out/r02n_ncu1.log gpurun_out/r02n_ncu2.log | cut -c1-90
