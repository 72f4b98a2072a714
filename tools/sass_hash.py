"""SASS fingerprints of kernels in librelion_b200.so, so that numbers taken from a committed ncu capture (DRAM traffic) are only
reported while the kernel they were measured on is unchanged.

    python tools/sass_hash.py record <report.ncu-rep> <workload> <pool> [out.json]   # traffic + fingerprints of the captured kernels
    python tools/sass_hash.py check  [traffic.json]                                   # which entries are still valid

The fingerprint is the SHA-1 of the instruction text of a function (`cuobjdump -sass -fun <mangled name>`, addresses and
encodings stripped)."""
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "relion_b200", "librelion_b200.so")
DEFAULT = os.path.join(ROOT, "profiles", "traffic_r02.json")


def _functions(text):
    """{mangled name: sha1 of its instruction text} from cuobjdump -sass output."""
    out, name, lines = {}, None, []
    def flush():
        if name is not None:
            out[name] = hashlib.sha1("\n".join(lines).encode()).hexdigest()
    for ln in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            flush()
            name, lines = m.group(1), []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", ln)
        if m and name is not None:
            lines.append(re.sub(r"\s+", " ", m.group(1)))
    flush()
    return out


def fingerprints(names=None, lib=LIB):
    cmd = ["cuobjdump", "-sass"]
    for n in names or []:
        cmd += ["-fun", n]
    txt = subprocess.run(cmd + [lib], capture_output=True, text=True).stdout
    return _functions(txt)


def base_name(kernel):
    return re.sub(r"^void ", "", kernel).split("<")[0].split("(")[0]


def record(rep, workload, pool, out_path=DEFAULT):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    col = {h: i for i, h in enumerate(rows[0])}
    unit = rows[1]
    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    allf = fingerprints()
    kernels = {}
    for r in rows[2:]:
        b = base_name(r[col["Kernel Name"]])
        rd = to_bytes(r[col["dram__bytes_read.sum"]], unit[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], unit[col["dram__bytes_write.sum"]])
        funcs = {n: h for n, h in allf.items() if b in n}
        kernels[b] = {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                      "duration_ms_under_ncu": float(r[col["gpu__time_duration.sum"]]), "functions": funcs}
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data[workload] = {"pool": int(pool), "source": os.path.basename(rep), "kernels": kernels}
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
    print("wrote", out_path, {k: v["dram_bytes"] for k, v in kernels.items()})


def traffic(workload, pool, kernel, path=DEFAULT):
    """DRAM bytes per launch of `kernel` (base name) from the committed capture, or None when the capture is of another
    workload / pool size or the kernel's SASS has changed since."""
    try:
        d = json.load(open(path)).get(workload)
        if not d or d.get("pool") != pool or kernel not in d["kernels"]:
            return None
        k = d["kernels"][kernel]
        now = fingerprints(list(k["functions"]))
        for name, sha in k["functions"].items():
            if now.get(name) != sha:
                return None
        return int(k["dram_bytes"])
    except Exception:
        return None


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "record":
        record(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else DEFAULT)
    elif len(sys.argv) >= 2 and sys.argv[1] == "check":
        path = sys.argv[2] if len(sys.argv) > 2 else DEFAULT
        for wl, d in json.load(open(path)).items():
            for k in d["kernels"]:
                print(wl, d["pool"], k, traffic(wl, d["pool"], k, path))
    else:
        print(__doc__)
