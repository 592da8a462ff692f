"""profiles/roofline_traffic.json from an `ncu --set full` capture of the latency kernel (bench.py reads it only when the library hash
matches the loaded libtaco2dec.so).  usage: python tools/make_roofline_traffic.py gpurun_out/<capture>.ncu-rep [source note]"""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tacotron2_subword_b200.build import source_stamp
rep = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]


def val(key):
    i = hdr.index(key)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


lib = os.path.join(ROOT, "tacotron2_subword_b200", "csrc", "libtaco2dec.so")
out = {
    "kernel": r[hdr.index("Kernel Name")][:120],
    "source": f"ncu --set full --clock-control none, {note}",
    "lib_sha256_16": source_stamp(),      # hash of the kernel sources (the .so itself is not byte-reproducible)
    "dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
    "dram_read_bytes": int(val("dram__bytes_read.sum")),
    "dram_write_bytes": int(val("dram__bytes_write.sum")),
    "l2_bytes_per_launch": int(val("l1tex__m_xbar2l1tex_read_bytes.sum")),       # bytes the SMs pulled through the crossbar (L2 -> SM)
    "l2_hit_rate_pct": val("lts__t_sector_hit_rate.pct"),
    "duration_ms_under_ncu": val("gpu__time_duration.sum") / 1e6 if "ns" in units[hdr.index("gpu__time_duration.sum")].lower() else val("gpu__time_duration.sum"),
}
json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
