for v in "" _u2 _u1 _hsqu1 _hsqu2 _hsqu4; do
  if [ -n "$v" ]; then export PHE_B200_LIB=$PWD/pailliercryptolib_python_b200/lib/libphe_b200$v.so; else unset PHE_B200_LIB; fi
  python bench.py --no-cpu --no-secondary --no-config5 --no-api > gpurun_out/r02_ab$v.json 2> gpurun_out/r02_ab$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r02_ab$v.json')); k=d['kernels']; print('$v', d['value'], d['ms_per_step'], k['k_dec_pair']['ms_total']/k['k_dec_pair']['launches'], k['k_encrypt_npair']['ms_total']/5)"
done
