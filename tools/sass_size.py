"""Static SASS instruction count per inlined source function (needs -lineinfo): python tools/sass_size.py <cubin>"""
import collections, re, subprocess, sys
dis = subprocess.run(["nvdisasm", "-g", "-c", sys.argv[1]], capture_output=True, text=True).stdout
impl = open("/root/repo/brax_tracking_b200/csrc/bt_impl.h").read().splitlines()
funcs = []
for n, l in enumerate(impl, 1):
    mm = re.match(r"  (?:static )?(?:template <[^>]*>\s*)?BT_DEV\s+[\w:<>\*&\s]+?\s+(\w+)\(", l)
    if mm: funcs.append((n, mm.group(1)))
def func_of(line):
    name = "?"
    for n, f in funcs:
        if n <= line: name = f
        else: break
    return name
cnt = collections.Counter(); cur = None
for l in dis.splitlines():
    if "//## File" in l:
        chain = re.findall(r'"([^"]*)", line (\d+)', l)
        fr = [(f.split("/")[-1], int(n)) for f, n in chain]
        impl_fr = [func_of(n) for f, n in fr if f == "bt_impl.h"]
        cur = impl_fr[-1] if impl_fr else (fr[-1][0] if fr else "?")
    elif re.match(r"\s*/\*[0-9a-f]+\*/", l):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot)
for k, v in cnt.most_common(30): print(f"{k:24s} {v:6d} {100*v/tot:5.1f}%")
