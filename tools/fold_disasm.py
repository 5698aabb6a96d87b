"""Analysis recipe (not product code): fold the -O0 disassembly produced by disasm_reference.py
into pseudo-C statements, one function at a time:  python fold_disasm.py _PassingIFtoIV
Used while writing oracle/pedfac_oracle.c; every oracle function cites the addresses it follows."""
import re, sys, json
names = {int(k,16):v for k,v in json.load(open('names.json')).items()}
lines = open('pedigraph.asm').read().split('\n')
funcs = {}; cur=None
for l in lines:
    m = re.match(r'<(\w+)>:', l)
    if m: cur=m.group(1); funcs[cur]=[]; continue
    if cur and l.strip(): funcs[cur].append(l)

R64 = ['rax','rbx','rcx','rdx','rsi','rdi','r8','r9','r10','r11','r12','r13','r14','r15','rbp','rsp']
alias = {}
for r in ['a','b','c','d']:
    for n in [f'r{r}x', f'e{r}x', f'{r}x', f'{r}l', f'{r}h']: alias[n]=f'r{r}x'
for r in ['si','di']:
    for n in [f'r{r}', f'e{r}', r, f'{r}l']: alias[n]=f'r{r}'
for i in range(8,16):
    for n in [f'r{i}', f'r{i}d', f'r{i}w', f'r{i}b']: alias[n]=f'r{i}'
for i in range(16): alias[f'xmm{i}']=f'xmm{i}'
alias['rbp']='rbp'; alias['rsp']='rsp'; alias['ebp']='rbp'

def fold(fn):
    body = funcs[fn]
    # jump targets
    targets=set()
    for l in body:
        m = re.match(r'([0-9a-f]+): (j\w+)\s+0x([0-9a-f]+)', l)
        if m: targets.add(int(m.group(3),16))
    reg={}
    out=[]
    def val(op):
        op=op.strip()
        m = re.match(r'(?:(QWORD|DWORD|WORD|BYTE|XMMWORD) PTR )?\[(.*)\]$', op)
        if m:
            inner=m.group(2)
            mm = re.match(r'rbp([+-]0x[0-9a-f]+)$', inner)
            if mm:
                off=int(mm.group(1),16)
                return ('v%x'%(-off)) if off<0 else ('arg%x'%off)
            if inner.startswith('rip'):
                return 'G[%s]'%inner
            # general: base + idx*scale + disp
            parts = re.split(r'(?=[+-])', inner)
            terms=[]
            for p in parts:
                sign = '-' if p.startswith('-') else '+'
                p=p.lstrip('+-')
                if '*' in p:
                    r,s=p.split('*'); terms.append((sign, '%s*%s'%(reg.get(alias.get(r,r),r), s)))
                elif p in alias:
                    terms.append((sign, reg.get(alias[p],p)))
                else: terms.append((sign,p))
            # pretty: base[idx] when scale matches
            s=''
            for i,(sg,t) in enumerate(terms):
                s += (t if i==0 and sg=='+' else f'{sg}{t}')
            return '*(%s)'%s
        if op in alias: return reg.get(alias[op], op)
        return op
    def setr(r, v):
        reg[alias.get(r,r)] = v
    for l in body:
        m = re.match(r'([0-9a-f]+): (\S+)\s*(.*?)(?:\s+; (.*))?$', l)
        if not m: continue
        a=int(m.group(1),16); mn=m.group(2); ops=m.group(3); note=m.group(4) or ''
        ops = re.sub(r'\s+#.*$','',ops)
        if a in targets:
            out.append('L_%x:'%a); reg={}
        o=[x.strip() for x in re.split(r',(?![^\[]*\])', ops)] if ops else []
        try:
            if mn in ('mov','movsd','movss','movsxd','movzx','movsx','movaps','movabs','movq','movd','cvtsi2sd','cvtsi2ss','cvtss2sd','cvtsd2ss','cvttsd2si','cvttss2si','movups','movapd'):
                src=val(o[1])
                if note.startswith('DBL'): src=note[4:]
                if mn.startswith('cvtsi2s'): src='(double)(%s)'%src
                if mn=='cvttsd2si' or mn=='cvttss2si': src='(int)(%s)'%src
                if mn=='cvtsd2ss': src='(float)(%s)'%src
                if o[0].startswith(('QWORD','DWORD','BYTE','WORD','XMMWORD','[')):
                    dst=val(o[0]); out.append('%x:  %s = %s'%(a,dst,src))
                else: setr(o[0],src)
            elif mn=='lea':
                mm=re.match(r'\[(.*)\]',o[1]); inner=mm.group(1)
                if inner.startswith('rip'):
                    setr(o[0], note if note else 'G[%s]'%inner)
                else:
                    v=val('['+inner+']'); setr(o[0], '&'+v[1:] if v.startswith('*') else '&'+v)
            elif mn in ('add','sub','imul','and','or','xor','shl','sar','shr','addsd','subsd','mulsd','divsd','addss','subss','mulss','divss','xorps','xorpd','andps','andpd','pxor'):
                sym={'add':'+','sub':'-','imul':'*','and':'&','or':'|','xor':'^','shl':'<<','sar':'>>','shr':'>>>','addsd':'+','subsd':'-','mulsd':'*','divsd':'/','addss':'+f','subss':'-f','mulss':'*f','divss':'/f','xorps':'^','xorpd':'^','andps':'&','andpd':'&','pxor':'^'}[mn]
                if len(o)==3: setr(o[0], '(%s%s%s)'%(val(o[1]),sym,val(o[2])))
                elif mn in('xor','xorps','xorpd','pxor') and o[0]==o[1]: setr(o[0],'0')
                else:
                    src=val(o[1])
                    if note.startswith('DBL'): src=note[4:]
                    if o[0].startswith(('QWORD','DWORD','BYTE','WORD','[')):
                        out.append('%x:  %s %s= %s'%(a,val(o[0]),sym,src))
                    elif alias.get(o[0])in('rsp',): pass
                    else: setr(o[0], '(%s%s%s)'%(val(o[0]),sym,src))
            elif mn in ('cmp','test','ucomisd','ucomiss','comisd'):
                src=val(o[1]); 
                if note.startswith('DBL'): src=note[4:]
                reg['flags']='%s(%s, %s)'%(mn,val(o[0]),src)
            elif mn.startswith('j') and mn!='jmp':
                out.append('%x:  if %s [%s] goto L_%s'%(a,reg.get('flags','?'),mn,o[0][2:]))
            elif mn=='jmp':
                out.append('%x:  goto L_%s'%(a,o[0][2:]))
            elif mn=='call':
                args=', '.join(str(reg.get(r,r)) for r in ['rdi','rsi','rdx','rcx','r8','r9'][:int(__import__('os').environ.get('NARGS','4'))])
                xa=', '.join(str(reg.get(r,'')) for r in ['xmm0','xmm1'] if r in reg)
                out.append('%x:  CALL %s(%s | %s)'%(a,note or ops,args,xa))
                for r in ['rax','rcx','rdx','rsi','rdi','r8','r9','r10','r11']: reg.pop(r,None)
                setr('rax','ret_%x'%a); setr('xmm0','fret_%x'%a)
            elif mn in ('cdq','cdqe','cqo','nop','push','pop','cs','data16'): pass
            elif mn=='idiv':
                d=val(o[0]); n=reg.get('rax','rax'); setr('rax','(%s / %s)'%(n,d)); setr('rdx','(%s %% %s)'%(n,d))
            elif mn=='neg': setr(o[0],'(-%s)'%val(o[0]))
            elif mn=='ret': out.append('%x:  return %s'%(a,reg.get('rax','')))
            elif mn.startswith('set'): setr(o[0], '%s[%s]'%(reg.get('flags','?'),mn))
            elif mn.startswith('cmov'): setr(o[0], '(%s[%s] ? %s : %s)'%(reg.get('flags','?'),mn,val(o[1]),val(o[0])))
            elif mn in ('sqrtsd','roundsd'): setr(o[0], '%s(%s)'%(mn, ','.join(val(x) for x in o[1:])))
            else: out.append('%x:  ASM %s %s %s'%(a,mn,ops,note))
        except Exception as e:
            out.append('%x:  ASM? %s %s %s (%s)'%(a,mn,ops,note,e))
    return out
if __name__=='__main__':
    for fn in sys.argv[1:]:
        print('=====',fn); print('\n'.join(fold(fn)))
