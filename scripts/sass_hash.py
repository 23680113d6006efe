"""Per-kernel SASS hashes of two builds of libmgicp.so: python scripts/sass_hash.py old.so new.so  (a refactor or an added kernel must leave
the verified kernels byte-identical; instruction encodings are compared, line-info comments are not)."""
import sys, re, hashlib, subprocess
def hashes(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d = {}; cur = None
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m: cur = m.group(1); d[cur] = hashlib.md5(); continue
        if cur and re.match(r"^\s+/\*[0-9a-f]{4}\*/", l):
            d[cur].update(re.sub(r"/\* 0x[0-9a-f]* \*/", "", l).encode())
    return {k: v.hexdigest()[:10] for k, v in d.items()}
a, b = hashes(sys.argv[1]), hashes(sys.argv[2])
for k in sorted(set(a) | set(b)):
    print(f"{k:40s} {a.get(k,'-'):12s} {b.get(k,'-'):12s} {'same' if a.get(k)==b.get(k) else 'DIFF'}")
