#!/bin/bash
# Round-2 baseline evidence with the round-1 build: OpenCL ICD probe, quick perf, tail probe, ncu full captures
# (config 3 and the 4K cut of config 4).  Every command runs once WITHOUT ncu first.
set -u
O=gpurun_out/r2_base; mkdir -p $O
{ echo "# ls /etc/OpenCL/vendors"; ls -la /etc/OpenCL/vendors 2>&1; echo "# ldconfig -p | grep -i opencl"; ldconfig -p | grep -i opencl;
  echo "# find nvidia opencl"; find / \( -name "*nvidia-opencl*" -o -name "nvidia.icd" -o -name "libnvidia-opencl*" \) 2>/dev/null | head;
  echo "# clinfo"; which clinfo 2>&1; echo "# nvidia-smi -L"; nvidia-smi -L; echo "# nproc"; nproc; } > $O/opencl_probe.txt 2>&1
python scripts/quick_perf.py synth > $O/quick_perf.txt 2>&1
RTG_LIB=$PWD/build_variants/librt_phase.so python scripts/tail_probe.py 4 > $O/tail_probe.txt 2>&1
python scripts/profile_case.py 256 3840 2160 1 6 2 > $O/c3_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c3 \
    python scripts/profile_case.py 256 3840 2160 1 6 1 > $O/ncu_c3.log 2>&1
python scripts/profile_case.py 1024 3840 2160 1 8 2 > $O/c4k_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c4k \
    python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_c4k.log 2>&1
cat $O/quick_perf.txt | cut -c1-200; cat $O/tail_probe.txt
