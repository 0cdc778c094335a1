for i in 1 2 3; do
  for lib in ab/libd4s_prev.so diffusions-for-sbi_b200/d4s/libd4s.so; do
    D4S_LIB_PATH=$PWD/$lib python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), round(d['e2e']['value']))"
  done
done
