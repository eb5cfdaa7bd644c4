#!/bin/bash
# Round-2 closing captures (run under gpurun; PFX=<prefix> ONLY="C5 C4" select name and configs): `ncu --set full` of the extend kernel on C5 / C4 / C1 with the final
# build, after a plain run of the same command; summarised on the box, the .ncu-rep deleted (64 MiB return limit).
set -x
LIB=go_raytracer_b200/csrc/libgrt_cuda.so
summ() {   # <rep> <name>
  python tools/summarize_profile.py gpurun_out/$1.ncu-rep > gpurun_out/$2.txt 2>&1
  python tools/profile_lines.py gpurun_out/$1.ncu-rep $LIB 70 > gpurun_out/$2_lines.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py >> gpurun_out/$2.txt
  rm -f gpurun_out/$1.ncu-rep
}
for cs in "C5 16 c5" "C4 64 c4" "C1 100 c1"; do
  set -- $cs
  case " ${ONLY:-C5 C4 C1} " in *" $1 "*) ;; *) continue;; esac
  Bc="python bench.py --config $1 --spp $2 --steps 1 --warmup 1 --no-cpu --no-configs"
  $Bc > gpurun_out/${PFX:-r2b}_plain_$1.json 2> gpurun_out/${PFX:-r2b}_plain_$1.err || continue
  ncu --set full --clock-control none --import-source on -k regex:wf_extend_dyn -s 40 -c 1 -f -o gpurun_out/prof_$3 $Bc > gpurun_out/${PFX:-r2b}_ncu_$3.log 2>&1
  summ prof_$3 ${PFX:-r2b}_$3_wf_extend_dyn
done
rm -f gpurun_out/${PFX:-r2b}_ncu_*.log
du -sh gpurun_out
