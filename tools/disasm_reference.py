"""Analysis recipe (not product code): regenerate the annotated disassembly of the reference's
Mach-O binary into ./pedigraph.asm + ./names.json. Run by hand in a scratch
directory in a container where /root/reference exists; nothing in tests/ or bench.py needs it."""
import struct, re, subprocess
d = open('/root/reference/exec/pedigraph','rb').read(); B = 0x100000000
symoff, nsyms, stroff = 235680, 602, 245684
names = {}
for i in range(nsyms):
    strx, typ, sect, desc, val = struct.unpack('<IBBHQ', d[symoff+16*i: symoff+16*i+16])
    if sect and (typ & 0xe) == 0xe:
        names[val] = d[stroff+strx: d.index(b'\0', stroff+strx)].decode()
ind = struct.unpack('<93I', d[245312:245312+93*4])
syms = [struct.unpack('<IBBHQ', d[symoff+16*i: symoff+16*i+16]) for i in range(nsyms)]
for k in range(45):
    strx = syms[ind[k]][0]
    names[0x10003028e+6*k] = d[stroff+strx: d.index(b'\0', stroff+strx)].decode() + '@stub'
cs, off = {}, 0x3056e
while off < 0x3056e+0x5553:
    e = d.index(b'\0', off)
    if e > off: cs[B+off] = d[off:e].decode(errors='replace')
    off = e+1
open('text.bin','wb').write(d[0xb10:0xb10+0x2f77d+0x10e+0x1d2])
asm = subprocess.run('objdump -D -b binary -m i386:x86-64 -M intel --adjust-vma=0x100000b10 text.bin',
                     shell=True, capture_output=True, text=True).stdout
out=[]
for line in asm.splitlines():
    m = re.match(r'\s+([0-9a-f]+):\t[^\t]*\t(.*)', line)
    if not m: continue
    a, ins = int(m.group(1),16), m.group(2)
    if a in names and '@stub' not in names[a]: out.append(f'\n<{names[a]}>:')
    t = re.search(r'(?:call|jmp)\s+0x([0-9a-f]+)|# 0x([0-9a-f]+)', ins)
    tgt = int(t.group(1) or t.group(2), 16) if t else None
    note = names.get(tgt) or (('STR '+repr(cs[tgt][:70])) if tgt in cs else
           (('DBL %r' % struct.unpack('<d', d[tgt-B:tgt-B+8])[0]) if tgt and 0x35ad0 <= tgt-B < 0x35f64 else ''))
    out.append(f'{a:x}: {ins}' + (f'   ; {note}' if note else ''))
open('pedigraph.asm','w').write('\n'.join(out))
import json
json.dump({hex(k):v for k,v in names.items()}, open('names.json','w'))
print(len(out))
print(struct.unpack('<27d', d[0x35b10:0x35b10+216]))
