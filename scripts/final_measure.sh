set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r2b.txt 2>&1; tail -3 gpurun_out/gpu_tests_r2b.txt
python bench.py > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err
python bench.py --impl reference > gpurun_out/bench_r2b_reference_arm.json 2> gpurun_out/bench_r2b_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2b.txt 2>&1; tail -2 gpurun_out/smoke_r2b.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-stereo --no-lba > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lba_r2b.csv python scripts/lba_prof.py > gpurun_out/ncu_lba_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ldlt_win|lin_edge|y_edge|lm_gather|rev_gather|sep_merge|schur_pairs" -c 9 -o gpurun_out/lba_r2b python scripts/lba_prof.py > gpurun_out/ncu_lba_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"resize_words|fast_warp" -c 3 -o gpurun_out/extract_r2b python scripts/profile_extract.py 64 1 --device > gpurun_out/ncu_extract_full.log 2>&1
ls -la gpurun_out/
