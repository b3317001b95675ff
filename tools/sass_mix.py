#!/usr/bin/env python3
"""Static SASS instruction mix per kernel of an object file (cuobjdump -sass), for profiles/r02_sass_excerpts.md.
usage: sass_mix.py file.o kernel_substring [kernel_substring ...]"""
import re
import subprocess
import sys


def main():
    sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
    cur, cnt = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = {}
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", line)
        if cur and m:
            cnt[cur][m.group(1)] = cnt[cur].get(m.group(1), 0) + 1
    for f, c in cnt.items():
        if not any(k in f for k in sys.argv[2:]):
            continue
        tot = sum(c.values())
        wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
        top = ", ".join("%s %d" % kv for kv in sorted(c.items(), key=lambda x: -x[1])[:8] if not kv[0].startswith("IMAD.WIDE"))
        print("| `%s` | %d | %d | %s |" % (subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()[:70], tot, wide, top))


if __name__ == "__main__":
    main()
