"""Attributes the SASS of one kernel to source lines: instruction counts and local-memory (spill / stack) instructions per
line.  Usage: sass_lines.py <object.o> <kernel name substring> [top]   (needs cuobjdump + nvdisasm; compile with -lineinfo)"""
import collections, os, re, subprocess, sys, tempfile

def main():
    obj, name = os.path.abspath(sys.argv[1]), sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", cub], cwd=d, check=True, capture_output=True, text=True).stdout
    cur, on = None, False
    tot, loc = collections.Counter(), collections.Counter()
    for l in txt.splitlines():
        if l.startswith(".text."):
            on = name in l
            continue
        if l.startswith(".section") and on and ".text." not in l:
            on = False
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", l):
            tot[cur] += 1
            if re.search(r"\b(STL|LDL)", l):
                loc[cur] += 1
    print("instructions", sum(tot.values()), "local-memory instructions", sum(loc.values()))
    byfile = collections.Counter()
    for k, v in tot.items():
        byfile[k[0] if k else None] += v
    print("by file:", dict(byfile))
    print("== local-memory instructions by line")
    for k, v in sorted(loc.items(), key=lambda x: -x[1])[:top]:
        print("  %4d  %s" % (v, k))
    print("== instructions by line")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        print("  %4d  %s" % (v, k))

main()
