"""Compare two `cuobjdump -sass` dumps function by function (instruction text only: addresses and encodings
stripped).  Used to prove that a source clean-up left the compiled hot kernels untouched -- the register-buffered
streaming loop is sensitive to ptxas scheduling, so "same SASS" is the cheap guarantee.
    cuobjdump -sass effort_b200/libeffort_b200.so > after.sass;  python tools/sass_diff.py before.sass after.sass"""
import re
import sys


def split(fn):
    funcs, cur = {}, None
    for line in open(fn, errors="replace"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);\s*/\*", line)
        if m:
            funcs[cur].append(re.sub(r"\s+", " ", m.group(1).strip()))
    return funcs


a, b = split(sys.argv[1]), split(sys.argv[2])
same = [f for f in a if f in b and a[f] == b[f]]
diff = [f for f in a if f in b and a[f] != b[f]]
print(f"identical: {len(same)}   changed: {len(diff)}   only in first: {len(set(a) - set(b))}   only in second: {len(set(b) - set(a))}")
for f in diff:
    print("  CHANGED", f[:100], len(a[f]), "->", len(b[f]), "instructions")
for f in sorted(set(a) - set(b)):
    print("  removed", f[:100])
for f in sorted(set(b) - set(a)):
    print("  added  ", f[:100])
sys.exit(1 if diff else 0)
