#!/bin/bash
# Round 2, session C: where does the event phase's time go?  (debug hooks build; results of these runs are WRONG by design)
TAG=${1:-r2c}
mkdir -p gpurun_out
OUT=gpurun_out/hooks_$TAG.txt
: > $OUT
export XMAPS_B200_LIB=$PWD/build_variants/libxm_dbg.so
for d in 0 16 17 18 19 80 81 48; do
  echo "== debug=$d" >> $OUT
  (timeout 120 python bench.py --quick --steps 10 --warmup 3 --opt debug=$d >> $OUT 2>> gpurun_out/hooks_$TAG.err)
done
echo "== debug=16 alive=0" >> $OUT
(timeout 120 python bench.py --quick --steps 10 --warmup 3 --opt debug=16 --opt alive=0 >> $OUT 2>> gpurun_out/hooks_$TAG.err)
echo "== debug=18 alive=0" >> $OUT
(timeout 120 python bench.py --quick --steps 10 --warmup 3 --opt debug=18 --opt alive=0 >> $OUT 2>> gpurun_out/hooks_$TAG.err)
echo "== debug=16 stage_xmap=0" >> $OUT
(timeout 120 python bench.py --quick --steps 10 --warmup 3 --opt debug=16 --opt stage_xmap=0 >> $OUT 2>> gpurun_out/hooks_$TAG.err)
unset XMAPS_B200_LIB
python - <<'PY'
import json
for ln in open('gpurun_out/hooks_r2c.txt'):
    ln=ln.strip()
    if ln.startswith('=='): print(ln, end='  ')
    elif ln.startswith('{'):
        d=json.loads(ln); print('frame_us %.2f' % d['frame_us'])
PY
tail -3 gpurun_out/hooks_$TAG.err
