mkdir -p gpurun_out/r02i /tmp/rep
cd /tmp/rep && cuobjdump -xelf all /root/repo/rav1d_b200/csrc/build/ipred.o > /dev/null 2>&1; ls /tmp/rep; cd /root/repo
timeout 500 ncu --set full --clock-control none --import-source on -o /tmp/rep/full_4k10c5 python tools/one_frame.py 4k10c5 2 > gpurun_out/r02i/ncu_full_c5.log 2>&1; echo ncufull_c5 rc=$?
ncu -i /tmp/rep/full_4k10c5.ncu-rep --page raw --csv > /tmp/rep/raw_4k10c5.csv 2>/dev/null
python tools/ncu_summary.py /tmp/rep/raw_4k10c5.csv > gpurun_out/r02i/ncu_full_summary_4k10c5.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:intra_levels -c 2 -o /tmp/rep/full_intra python tools/one_intra_frame.py > gpurun_out/r02i/ncu_intra.log 2>&1; echo ncu_intra rc=$?
ncu -i /tmp/rep/full_intra.ncu-rep --page raw --csv > /tmp/rep/raw_intra.csv 2>/dev/null
python tools/ncu_summary.py /tmp/rep/raw_intra.csv > gpurun_out/r02i/ncu_full_summary_intra.csv
ncu -i /tmp/rep/full_intra.ncu-rep --page source --csv --print-source sass > /tmp/rep/sass_intra.csv 2>/dev/null
python tools/ncu_lines.py /tmp/rep/sass_intra.csv /tmp/rep/ipred.sm_100a.cubin "intra_levels_kernelINS_4BD16" 80 > gpurun_out/r02i/intra_lines.txt 2>&1
ncu -i /tmp/rep/full_intra.ncu-rep --page details 2>/dev/null | head -400 > gpurun_out/r02i/details_intra.txt
find gpurun_out -size +8M -delete
du -sh gpurun_out; ls -la gpurun_out/r02i
