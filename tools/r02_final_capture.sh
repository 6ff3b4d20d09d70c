# Round-2 final measurements on one B200 (run under gpurun): GPU tests, the bench line, the ncu launch list of the bench
# command and one full ncu capture of the dominant kernel.  Numbers printed under ncu are never bench values.
python -m pytest tests -m gpu -x -q > gpurun_out/r02_final5_pytest.log 2>&1; tail -3 gpurun_out/r02_final5_pytest.log
python bench.py > gpurun_out/r02_final5_n1.json 2> gpurun_out/r02_final5_n1.err; tail -c 600 gpurun_out/r02_final5_n1.err
python bench.py --impl reference > gpurun_out/r02_final5_ref.json 2> gpurun_out/r02_final5_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches4_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-api --no-secondary --no-config5 > gpurun_out/r02_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dec_pair -s 1 -c 1 -f -o gpurun_out/r02_dec2048_final5 python tools/tail_probe.py 100000 > gpurun_out/r02_ncu_dec.log 2>&1; tail -2 gpurun_out/r02_ncu_dec.log
ls -la gpurun_out/*.ncu-rep
