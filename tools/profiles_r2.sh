#!/bin/bash
# Round-2 evidence batch (run under gpurun): launch lists of the bench commands and one `ncu --set full` capture per
# dominant kernel, each only after its own command has exited 0 without ncu.  The .ncu-rep files are summarised ON the
# box (tools/summarize_profile.py, tools/profile_lines.py) and deleted: gpurun brings back at most 64 MiB.
set -x
LIB=go_raytracer_b200/csrc/libgrt_cuda.so
summ() {   # <rep> <name>
  python tools/summarize_profile.py gpurun_out/$1.ncu-rep > gpurun_out/$2.txt 2>&1
  python tools/profile_lines.py gpurun_out/$1.ncu-rep $LIB 45 > gpurun_out/$2_lines.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py >> gpurun_out/$2.txt
  rm -f gpurun_out/$1.ncu-rep
}
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-configs"
$B > gpurun_out/r2_plain_c2.json 2> gpurun_out/r2_plain_c2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c2.csv $B > gpurun_out/r2_ncu_ll_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_mega -s 3 -c 1 -f -o gpurun_out/prof_mega python bench.py --steps 1 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_ncu_full_c2.log 2>&1
summ prof_mega r2_mega_fullsize
for cs in "C4 16" "C5 4" "C1 100" "C3 64"; do
  set -- $cs; c=$1; spp=$2
  Bc="python bench.py --config $c --spp $spp --steps 1 --warmup 1 --no-cpu --no-configs"
  $Bc > gpurun_out/r2_plain_$c.json 2> gpurun_out/r2_plain_$c.err || continue
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_$c.csv $Bc > gpurun_out/r2_ncu_ll_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:wf_extend_dyn -s 40 -c 1 -f -o gpurun_out/prof_c5 python bench.py --config C5 --spp 16 --steps 1 --warmup 1 --no-cpu --no-configs > gpurun_out/r2_ncu_full_c5.log 2>&1
summ prof_c5 r2_c5_wf_extend_dyn
ncu --set full --clock-control none --import-source on -k regex:wf_extend_dyn -s 40 -c 1 -f -o gpurun_out/prof_c4 python bench.py --config C4 --spp 64 --steps 1 --warmup 1 --no-cpu --no-configs > gpurun_out/r2_ncu_full_c4.log 2>&1
summ prof_c4 r2_c4_wf_extend_dyn
rm -f gpurun_out/r2_ncu_ll_*.log gpurun_out/r2_ncu_full_*.log
du -sh gpurun_out
