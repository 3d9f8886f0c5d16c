import re, sys, collections
f, s, e = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
full = len(sys.argv) > 4
cnt = collections.Counter()
for ln in open(f):
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', ln)
    if not m: continue
    a = int(m.group(1), 16)
    if not (s <= a < e): continue
    w = m.group(2).split()
    op = w[1] if w[0].startswith('@') else w[0]
    if not full: op = op.split('.')[0] + ('.16x2' if '16x2' in op else '')
    cnt[op] += 1
for k, v in cnt.most_common(): print(f"{v:5d} {k}")
print(f"{sum(cnt.values()):5d} TOTAL")
