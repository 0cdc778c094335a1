# same-box A/B of two builds of libd4s.so: usage  bash tests/gpu_ab.sh <task> <libA> <libB>
task=${1:-slcp}; A=${2:-ab/libd4s_base.so}; B=${3:-diffusions-for-sbi_b200/d4s/libd4s.so}
for i in 1 2 3; do
  for lib in $A $B; do
    D4S_LIB_PATH=$PWD/$lib python bench.py --task $task --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$task $lib', round(d['value']), round(d['e2e']['value']))"
  done
done
