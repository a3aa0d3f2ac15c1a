#!/bin/bash
# round 2, last GPU call on one GPU: smoke(), and a full ncu capture of ONE cfg4 launch of the final pair kernel
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none -k regex:'w8a8_gemm2' -c 2 -o gpurun_out/r3e_gemm_full -f python tools/ncu_gemm.py > gpurun_out/r3e_ncu.log 2>&1
tail -2 gpurun_out/r3e_ncu.log
ncu -i gpurun_out/r3e_gemm_full.ncu-rep --page raw --csv > gpurun_out/r3e_gemm_raw.csv 2>/dev/null
ls -la gpurun_out/r3e_gemm_full.ncu-rep; rm -f gpurun_out/r3e_gemm_full.ncu-rep
python tools/ncu_one_kernel_md.py gpurun_out/r3e_gemm_raw.csv "x" | head -30
