mkdir -p gpurun_out
python tools/profile_program.py --batch 64 --top 5 --out gpurun_out/prof64.json > gpurun_out/prof64.log 2>&1
python tools/tc_selftest.py --group ru --batch 64 > gpurun_out/ru_dbg0.log 2>&1
B2C_TC_DEBUG=1 python tools/tc_selftest.py --group ru --batch 64 > gpurun_out/ru_dbg1.log 2>&1
python tools/tc_selftest.py --group enc --batch 64 > gpurun_out/enc_dbg0.log 2>&1
B2C_TC_DEBUG=1 python tools/tc_selftest.py --group enc --batch 64 > gpurun_out/enc_dbg1.log 2>&1
python tools/tc_selftest.py --group dec --batch 64 --precs bf16 > gpurun_out/dec_dbg0.log 2>&1
B2C_TC_DEBUG=1 python tools/tc_selftest.py --group dec --batch 64 --precs bf16 > gpurun_out/dec_dbg1.log 2>&1
tail -3 gpurun_out/prof64.log
