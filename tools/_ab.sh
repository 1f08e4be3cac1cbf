cd "$(dirname "$0")/.."
for v in "$@"; do
  cp slam-robot_b200/csrc/libslamfe_$v.so slam-robot_b200/csrc/libslamfe.so
  touch slam-robot_b200/csrc/libslamfe.so
  echo "== $v"
  python bench.py --steps 5 --warmup 3 --no-cpu --no-other 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['phase_ms'])"
done
