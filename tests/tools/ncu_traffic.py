"""DRAM bytes per launch of the march kernels, from an `ncu --set full` report of the bench command -> the
profiles/rNN_traffic.json that bench.py quotes as `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum).
Dev tool, run on the build box on a report brought back from the GPU box:

    gpurun -- 'ncu --set full --clock-control none --import-source on -k regex:march_.*quad -c 4 \
               -o gpurun_out/r02_c3 python bench.py --steps 1 --warmup 3 --skip-extras'
    python tests/tools/ncu_traffic.py gpurun_out/r02_c3.ncu-rep profiles/r02_traffic.json "C3, 2^20 rays, <build>"

The LAST launch of each kernel in the report is taken (earlier ones are warm-up steps)."""
import csv, io, json, subprocess, sys

rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"].split("(")[0].split("<")[0].replace("svoxb::", "").replace("void ", "").strip()
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[m].replace(",", "")) * scale[u[m]]
    res[name] = tot
res["_source"] = f"{rep}: dram__bytes_read.sum + dram__bytes_write.sum of the last launch of each kernel; {what}"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
