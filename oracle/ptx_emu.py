"""
ORACLE (test infrastructure only) -- a small PTX interpreter.

The reference keeps three pieces of the accumulation path as inline PTX inside
cuburn/code/iter.py and cuburn/code/util.py: `trunca` (util.py:194-200), the packed-cell
add / overflow spill of the iterate kernel (iter.py:332-407) and the unpack + hotspot
flag computation of `flush_atom` (iter.py:429-540).  oracle/build_ref.py lifts those
texts verbatim into oracle/_ref/ref_kernels.json; this module executes them, lane by
lane semantics on numpy vectors, so that the oracle's and the device's packed-cell
arithmetic can be pinned against the reference's own instructions.

Supported: the subset those blocks use -- .reg declarations, predication (@p / @!p),
forward branches to a label, integer / float / bit-field arithmetic, cvt, setp (with a
boolean combine), mov with {lo, hi} packing, vote.ballot (per 32 lanes), global loads /
stores / red / atom on byte-addressed regions, and suld through a callback.
Floating point: .rn with round-to-nearest-even float32 results; .ftz is honoured for
results (subnormals flush to zero), which none of the pinned vectors reach.
"""
import re

import numpy as np

_TYPES = {'u32': np.uint32, 'b32': np.uint32, 's32': np.int32, 'f32': np.float32,
          'u64': np.uint64, 'b64': np.uint64, 'pred': np.bool_}


class Memory(object):
    """Byte-addressed regions: name -> (base address, numpy uint8 array)."""
    def __init__(self):
        self.regions = []

    def add(self, base, array):
        raw = array.view(np.uint8).reshape(-1)
        self.regions.append((int(base), raw))
        return int(base)

    def _find(self, addr, size):
        for base, raw in self.regions:
            if base <= addr and addr + size <= base + raw.size:
                return raw, addr - base
        raise IndexError('address %#x (+%d) is outside every region' % (addr, size))

    def load(self, addr, dtype, count=1):
        size = np.dtype(dtype).itemsize * count
        raw, off = self._find(int(addr), size)
        return raw[off:off + size].view(dtype).copy()

    def store(self, addr, dtype, values):
        v = np.asarray(values, dtype).reshape(-1)
        raw, off = self._find(int(addr), v.nbytes)
        raw[off:off + v.nbytes] = v.view(np.uint8)


def _rni(f, signed):
    f = np.asarray(f, np.float32).astype(np.float64)
    r = np.rint(f)
    lo, hi = (-2147483648.0, 2147483647.0) if signed else (0.0, 4294967295.0)
    r = np.where(np.isnan(r), 0.0, np.clip(r, lo, hi))
    return r.astype(np.int64).astype(np.int32 if signed else np.uint32)


def _ftz(x):
    x = np.asarray(x, np.float32)
    tiny = (np.abs(x) < np.float32(1.17549435e-38)) & (x != 0)
    return np.where(tiny, np.float32(0.0) * np.sign(x), x).astype(np.float32)


class Program(object):
    """Statements (split at ';') and label positions of one inline-asm block."""
    def __init__(self, text):
        text = re.sub(r'//[^\n]*', '', text)
        self.stmts, self.labels = [], {}
        buf = ''

        def flush():
            nonlocal buf
            for st in buf.split(';'):
                st = ' '.join(st.split())
                if st and st not in ('{', '}'):
                    self.stmts.append(st)
            buf = ''
        for line in text.split('\n'):
            stripped = line.strip()
            m = re.fullmatch(r'([A-Za-z_]\w*):', stripped)
            if m:
                flush()
                self.labels[m.group(1)] = len(self.stmts)
            elif stripped in ('{', '}'):
                flush()
            else:
                buf += ' ' + line
        flush()


class Machine(object):
    """Executes a Program for `lanes` threads at once."""
    def __init__(self, program, lanes, params, memory, special=None, suld=None):
        self.p, self.n, self.mem = program, lanes, memory
        self.params = params            # list: %0, %1, ... (scalars or per-lane arrays)
        self.special = special or {}
        self.suld = suld
        self.regs, self.types = {}, {}

    # ---- operands ------------------------------------------------------------------
    def _bcast(self, v, dtype):
        a = np.asarray(v)
        if a.ndim == 0:
            a = np.full(self.n, a)
        return a.astype(dtype)

    def val(self, tok, dtype):
        tok = tok.strip()
        if tok in self.regs:
            v = self.regs[tok]
            if v.dtype == dtype:
                return v
            if v.dtype.itemsize == np.dtype(dtype).itemsize:
                return v.view(dtype)
            return v.astype(dtype)
        if re.fullmatch(r'%\d+', tok):
            v = self._bcast(self.params[int(tok[1:])], np.asarray(self.params[int(tok[1:])]).dtype)
            if v.dtype.kind == 'f' or np.dtype(dtype).kind == 'f':
                return v.astype(dtype) if v.dtype.kind == np.dtype(dtype).kind else v.view(dtype)
            return v.astype(dtype)
        if tok in self.special:
            return self._bcast(self.special[tok], dtype)
        # immediates: numbers or parenthesised constant expressions
        expr = tok
        value = eval(expr, {'__builtins__': {}})
        return self._bcast(value, dtype)

    def set(self, name, value, mask):
        name = name.strip()
        dtype = self.types[name]
        value = np.asarray(value).astype(dtype) if np.asarray(value).dtype != dtype else value
        cur = self.regs[name]
        self.regs[name] = np.where(mask, value, cur).astype(dtype)

    # ---- execution -----------------------------------------------------------------
    def run(self):
        n = self.n
        waiting = {}                    # label -> lanes parked until that label
        active = np.ones(n, bool)
        pc = 0
        label_at = {}
        for name, idx in self.p.labels.items():
            label_at.setdefault(idx, []).append(name)
        while pc <= len(self.p.stmts):
            for name in label_at.get(pc, []):
                if name in waiting:
                    active = active | waiting.pop(name)
            if pc == len(self.p.stmts):
                break
            st = self.p.stmts[pc]
            pc += 1
            if st.startswith('.reg'):
                m = re.match(r'\.reg\s+\.(\w+)\s+(.*)', st, re.S)
                dtype = _TYPES[m.group(1)]
                for name in m.group(2).split(','):
                    name = name.strip()
                    self.regs[name] = np.zeros(n, dtype)
                    self.types[name] = dtype
                continue
            mask = active
            m = re.match(r'@(!?)(\w+)\s+(.*)', st, re.S)
            if m:
                pred = self.regs[m.group(2)]
                mask = active & (~pred if m.group(1) else pred)
                st = m.group(3)
            op, _, rest = st.partition(' ')
            args = self._split(rest)
            if op == 'bra':
                label = args[0]
                waiting[label] = waiting.get(label, np.zeros(n, bool)) | mask
                active = active & ~mask
                continue
            self._exec(op, args, mask)
        return self.regs

    @staticmethod
    def _split(rest):
        out, depth, cur = [], 0, ''
        for ch in rest:
            if ch in '{[(':
                depth += 1
            elif ch in '}])':
                depth -= 1
            if ch == ',' and depth == 0:
                out.append(cur.strip())
                cur = ''
            else:
                cur += ch
        if cur.strip():
            out.append(cur.strip())
        return out

    def _addr(self, tok):
        m = re.fullmatch(r'\[(\w+)(?:\s*\+\s*(\d+))?\]', tok.strip())
        base = self.val(m.group(1), np.uint64)
        return base + np.uint64(int(m.group(2) or 0))

    def _exec(self, op, a, mask):
        parts = op.split('.')
        root = parts[0]
        n = self.n
        if root in ('fma', 'mul', 'add') and parts[-1] == 'f32':
            x, y = self.val(a[1], np.float32).astype(np.float64), self.val(a[2], np.float32).astype(np.float64)
            if root == 'fma':
                r = x * y + self.val(a[3], np.float32).astype(np.float64)
            elif root == 'mul':
                r = x * y
            else:
                r = x + y
            self.set(a[0], _ftz(r.astype(np.float32)), mask)
        elif root == 'add' and parts[-1] == 'u64':
            self.set(a[0], self.val(a[1], np.uint64) + self.val(a[2], np.uint64), mask)
        elif root == 'cvt':
            dst, src = parts[-2], parts[-1]
            if src == 'f32' and dst in ('u32', 's32'):
                assert 'rni' in parts
                self.set(a[0], _rni(self.val(a[1], np.float32), dst == 's32').view(np.uint32), mask)
            elif dst == 'f32':
                self.set(a[0], self.val(a[1], _TYPES[src]).astype(np.float32), mask)
            else:
                self.set(a[0], self.val(a[1], _TYPES[src]).astype(_TYPES[dst]), mask)
        elif root == 'shl':
            s = self.val(a[2], np.uint32)
            v = self.val(a[1], np.uint32).astype(np.uint64) << np.minimum(s, 32).astype(np.uint64)
            self.set(a[0], (v & np.uint64(0xffffffff)).astype(np.uint32), mask)
        elif root == 'shr':
            s = np.minimum(self.val(a[2], np.uint32), 32).astype(np.uint64)
            self.set(a[0], (self.val(a[1], np.uint32).astype(np.uint64) >> s).astype(np.uint32), mask)
        elif root == 'and':
            self.set(a[0], self.val(a[1], np.uint32) & self.val(a[2], np.uint32), mask)
        elif root == 'bfe':
            v, pos, ln = (self.val(t, np.uint32).astype(np.uint64) for t in a[1:4])
            self.set(a[0], ((v >> pos) & ((np.uint64(1) << ln) - np.uint64(1))).astype(np.uint32), mask)
        elif root == 'bfi':
            src, base, pos, ln = (self.val(t, np.uint32).astype(np.uint64) for t in a[1:5])
            fld = ((np.uint64(1) << ln) - np.uint64(1)) << pos
            r = (base & ~fld & np.uint64(0xffffffff)) | ((src << pos) & fld)
            self.set(a[0], r.astype(np.uint32), mask)
        elif root == 'mov':
            if a[0].startswith('{'):            # unpack
                lo, hi = [t.strip() for t in a[0][1:-1].split(',')]
                v = self.val(a[1], np.uint64)
                self.set(lo, (v & np.uint64(0xffffffff)).astype(np.uint32), mask)
                self.set(hi, (v >> np.uint64(32)).astype(np.uint32), mask)
            elif a[1].startswith('{'):          # pack
                lo, hi = [t.strip() for t in a[1][1:-1].split(',')]
                v = self.val(lo, np.uint32).astype(np.uint64) | \
                    (self.val(hi, np.uint32).astype(np.uint64) << np.uint64(32))
                self.set(a[0], v, mask)
            else:
                self.set(a[0], self.val(a[1], self.types[a[0].strip()]), mask)
        elif root == 'setp':
            cmp_ = parts[-2] if parts[1] != 'and' else parts[2]
            dtype = _TYPES[parts[-1]]
            x, y = self.val(a[1], dtype), self.val(a[2], dtype)
            r = {'le': x <= y, 'lo': x < y, 'lt': x < y, 'eq': x == y, 'gt': x > y,
                 'ge': x >= y, 'hi': x > y, 'ne': x != y}[cmp_]
            if parts[1] == 'and':
                r = r & self.regs[a[3].strip()]
            self.set(a[0], r, mask)
        elif root == 'vote':
            pred = self.regs[a[1].strip()] & mask
            out = np.zeros(n, np.uint32)
            for w0 in range(0, n, 32):
                bits = 0
                for l in range(min(32, n - w0)):
                    if pred[w0 + l]:
                        bits |= 1 << l
                out[w0:w0 + 32] = bits
            self.set(a[0], out, mask)
        elif root == 'suld':
            lo, hi = [t.strip() for t in a[0][1:-1].split(',')]
            m = re.fullmatch(r'\[(\w+)\s*$', a[1].strip()) or re.match(r'\[(\w+)', a[1])
            coords = re.search(r'\{(.*?)\}', ','.join(a[1:])).group(1).split(',')
            xb, y = self.val(coords[0], np.uint32), self.val(coords[1], np.uint32)
            vlo, vhi = self.suld(xb, y)
            self.set(lo, vlo, mask)
            self.set(hi, vhi, mask)
        elif root in ('red', 'atom'):
            self._atomic(root, parts, a, mask)
        elif root == 'ld':
            addr = self._addr(a[1])
            if 'v2' in parts or 'v4' in parts:
                names = [t.strip() for t in a[0][1:-1].split(',')]
                dtype = _TYPES[parts[-1]]
                vals = np.zeros((n, len(names)), dtype)
                for l in np.flatnonzero(mask):
                    vals[l] = self.mem.load(addr[l], dtype, len(names))
                for k, name in enumerate(names):
                    self.set(name, vals[:, k], mask)
            else:
                dtype = _TYPES[parts[-1]]
                vals = np.zeros(n, dtype)
                for l in np.flatnonzero(mask):
                    vals[l] = self.mem.load(addr[l], dtype)[0]
                self.set(a[0], vals, mask)
        elif root == 'st':
            addr = self._addr(a[0])
            if 'v4' in parts or 'v2' in parts:
                names = [t.strip() for t in a[1][1:-1].split(',')]
                dtype = _TYPES[parts[-1]]
                cols = [self.val(t, dtype) for t in names]
                for l in np.flatnonzero(mask):
                    self.mem.store(addr[l], dtype, [c[l] for c in cols])
            else:
                dtype = _TYPES[parts[-1]]
                v = self.val(a[1], dtype)
                for l in np.flatnonzero(mask):
                    self.mem.store(addr[l], dtype, v[l])
        else:
            raise NotImplementedError('PTX instruction %s' % op)

    def _atomic(self, root, parts, a, mask):
        kind, dtype = parts[-2], _TYPES[parts[-1]]
        if root == 'red':
            addr, v = self._addr(a[0]), self.val(a[1], dtype)
            for l in np.flatnonzero(mask):
                cur = self.mem.load(addr[l], dtype)[0]
                with np.errstate(over='ignore'):
                    self.mem.store(addr[l], dtype, dtype(cur + v[l]))
            return
        addr, v = self._addr(a[1]), self.val(a[2], dtype)
        old = np.zeros(self.n, dtype)
        for l in np.flatnonzero(mask):
            cur = self.mem.load(addr[l], dtype)[0]
            old[l] = cur
            with np.errstate(over='ignore'):
                self.mem.store(addr[l], dtype, dtype(cur + v[l]) if kind == 'add' else v[l])
        self.set(a[0], old, mask)


def run(text, lanes, params, memory, special=None, suld=None):
    return Machine(Program(text), lanes, params, memory, special, suld).run()
