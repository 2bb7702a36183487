python profiles/assembly_phases.py C5 2>&1 | tail -4
nsys_missing=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/asm_launches.csv python profiles/assembly_phases.py C5 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/asm_launches.csv "assembly_phases.py C5" | head -30
