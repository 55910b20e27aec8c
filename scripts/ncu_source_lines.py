"""Per-source-line view of one kernel from an .ncu-rep (`ncu --set full --import-source on`, compiled with -lineinfo):
stall samples and executed warp instructions per CUDA source line.

    python scripts/ncu_source_lines.py gpurun_out/x.ncu-rep [kernel-substring] [top N]
"""
import csv
import subprocess
import sys


def main(path, kernel="", top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    out, file, func, seen_funcs = {}, None, None, []
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            if func not in seen_funcs:
                seen_funcs.append(func)
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or kernel not in (func or ""):
            continue
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try:
            key = (file, int(r[0]))
            samples, inst = int(r[i_s]), int(r[i_i])
        except ValueError:
            continue
        stalls = {h: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i])}
        e = out.setdefault(key, {"samples": 0, "inst": 0, "src": r[1], "stalls": {}})
        e["samples"] += samples
        e["inst"] += inst
        for k, v in stalls.items():
            e["stalls"][k] = e["stalls"].get(k, 0) + v
    total_s = sum(e["samples"] for e in out.values()) or 1
    total_i = sum(e["inst"] for e in out.values()) or 1
    print(f"# kernel filter '{kernel}': {total_s} samples, {total_i} warp instructions (all launches in the report)")
    print("file:line,samples_pct,inst_pct,top_stalls,source")
    for key, e in sorted(out.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = " ".join(f"{k[6:]}={v}" for k, v in sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print(f"{key[0]}:{key[1]},{100 * e['samples'] / total_s:.1f},{100 * e['inst'] / total_i:.1f},{st},\"{e['src'].strip()[:100]}\"")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 40)
