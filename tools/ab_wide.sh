for v in "" _g10 _g20 _g30 _u2 _u5g15; do
  if [ -n "$v" ]; then export PHE_B200_LIB=$PWD/pailliercryptolib_python_b200/lib/libphe_b200$v.so; else unset PHE_B200_LIB; fi
  echo "variant[$v] $(python tools/config5_3072.py --count 100000 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.readline()); print(round(d["ms_decrypt"],2), round(d["ms_encrypt"],2))')"
done
