#!/usr/bin/env python
"""Per-kernel SASS comparison of two builds of libllama3_b200.so (names normalised): which kernels were added,
removed or changed.  Used to prove that a change left the hardware-validated kernels byte-identical."""
import subprocess, re, hashlib, sys
def funcs(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d = {}; name = None; h = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name: d[name] = h.hexdigest()
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__X_", m.group(1))
            name = re.sub(r"_cu_[0-9a-f]{8}", "_cu_X", name)
            h = hashlib.md5(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\*", line)
        if m and h is not None: h.update(m.group(1).encode())
    if name: d[name] = h.hexdigest()
    return d
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
print(len(a), len(b))
print("only in new:", [k[:90] for k in b if k not in a])
print("only in old:", [k[:90] for k in a if k not in b])
print("changed:", [k[:90] for k in a if k in b and a[k] != b[k]])
