set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/nu_launches.csv python tools/bench_configs.py --only "C3-wide neutra_hmc funnel d=100 n=262144 H=256" > gpurun_out/nu_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:neutra_unwind -s 30 -c 1 -o gpurun_out/prof_nu_r02 -f python tools/bench_configs.py --only "C3-wide neutra_hmc funnel d=100 n=262144 H=256" > gpurun_out/nu_ncu2.log 2>&1
ls -la gpurun_out/prof_nu_r02.ncu-rep
