for d in 0 1 2 4 6 7; do
  echo "== B2C_TC_DEBUG=$d (1 skip epilogue work, 2 no TMA, 4 no MMA)"
  B2C_TC_DEBUG=$d timeout 200 python tools/power_probe.py --only "enc1,k7 C256,k1 C256,dec4" --no-program --secs 2.0 2>&1 | grep -v Warn | tail -5
done
