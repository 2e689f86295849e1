"""Dump one kernel's SASS from libtab200.so and count opcodes inside its hottest loop
(dev tool): python tools/sass_loop.py <mangled-name-substring>"""
import collections, re, subprocess, sys
so = 'tensoralloy_b200/csrc/libtab200.so'
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
blocks = re.split(r'\n\s*Function : ', txt)
for b in blocks[1:]:
    name = b.split('\n', 1)[0]
    if sys.argv[1] not in name:
        continue
    ins = []
    for line in b.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    # backward branches = loops; take the one with the largest body
    best = None
    for addr, text in ins:
        m = re.search(r'BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)', text)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            if best is None or addr - tgt > best[1] - best[0]:
                best = (tgt, addr)
    print(name[:90], 'instructions', len(ins), 'loop', best)
    cnt = collections.Counter()
    for addr, text in ins:
        if best and best[0] <= addr <= best[1]:
            op = text.split()[1] if text.startswith('@') else text.split()[0]
            cnt[op.split('.')[0]] += 1
    tot = sum(cnt.values())
    fp64 = sum(v for k, v in cnt.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
    print('  loop total', tot, 'fp64', fp64, dict(cnt.most_common(14)))
