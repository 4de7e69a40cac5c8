mkdir -p gpurun_out
for PAD in 0 4352 1052672 16384 147456; do
  echo "== PAD $PAD"
  B2C_ARENA_PAD=$PAD python tools/profile_program.py --batch 64 --top 4 2>&1 | tail -6
done
