# usage: bash tools/ab_lib.sh <variant suffix> ...   (libphe_b200_<suffix>.so in lib/); prints headline + config5 decrypt per variant
for v in "" "$@" "" "$@"; do
  if [ -n "$v" ]; then export PHE_B200_LIB=$PWD/pailliercryptolib_python_b200/lib/libphe_b200_$v.so; else unset PHE_B200_LIB; fi
  python bench.py --no-cpu --no-secondary --no-api > gpurun_out/ab_lib.json 2> gpurun_out/ab_lib.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_lib.json')); k=d['kernels']; print('variant[$v]', round(d['value']), round(d['ms_per_step'],2), 'dec', round(k['k_dec_pair']['ms_total']/5,2), 'c5 dec', round(d['config5']['ms_decrypt'],2))"
done
