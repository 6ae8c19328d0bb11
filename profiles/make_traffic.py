"""profiles/r2_traffic.json from an `ncu --set full` capture of one step (scripts/step_once.py):

    ncu -i gpurun_out/r2_step_prof.ncu-rep --page raw --csv > /tmp/step_raw.csv
    python profiles/make_traffic.py /tmp/step_raw.csv 3670012 > profiles/r2_traffic.json

DRAM bytes per launch (read + write) of the four in-step kernels; bench.py reports them as
`roofline*.traffic` for the same workload (ncu cannot run inside the bench).  ncu flushes the caches
before every launch (cold L2) and leaves dirty lines of the launch in L2: the write side of a
write-heavy kernel is undercounted (assembly: 155 MB are written, ~100 MB reach DRAM within the launch)."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(hdr)}
names = {"assemble_tiles": "assembly", "tree_factor_solve": "tree", "edge_backsub": "backsub", "spmv_pipe": "residual"}
out = {"n_dofs_per_gpu": int(sys.argv[2]), "source": "ncu --set full --clock-control none, one step of scripts/step_once.py 20 (cold L2 per launch)"}
for r in rows[2:]:
    for key, short in names.items():
        if key in r[ix["Kernel Name"]] and short not in out:
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[units[ix["dram__bytes_read.sum"]]]
            rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * scale
            wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[units[ix["dram__bytes_write.sum"]]]
            tscale = {"us": 1e-3, "ms": 1.0, "ns": 1e-6}[units[ix["gpu__time_duration.sum"]]]
            out[short] = {"kernel": r[ix["Kernel Name"]].split("(")[0], "dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                          "ms": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * tscale}
print(json.dumps(out, indent=1))
