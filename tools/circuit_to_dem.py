#!/usr/bin/env python3
"""Restricted stim-circuit -> detector-error-model analyser (fixture generator; not product code).

The reference's integration tests turn each `test_data/*.stim` circuit into a DEM with
stim's ErrorAnalyzer (decompose_errors=False), see
/root/reference/src/chromobius/decode/decoder_integration.test.cc:24-29,48-49.
stim is not installed in this image, so this script restates that conversion for the gate
set that occurs in the reference fixtures:

    QUBIT_COORDS R RX M MX MPP CX TICK X_ERROR Z_ERROR DEPOLARIZE1 DEPOLARIZE2
    DETECTOR OBSERVABLE_INCLUDE SHIFT_COORDS REPEAT

Method: Pauli-frame sensitivity sets propagated backwards through the circuit.
For every qubit q we keep xs[q] / zs[q] = the set of detectors+observables flipped by an
X / Z error on q at the current point in time.  Going backwards:

  M q      (record m)   xs[q] ^= S(m)            S(m) = detectors/observables containing m
  MX q                  zs[q] ^= S(m)
  MPP P1*P2..           X factor: zs ^= S(m); Z factor: xs ^= S(m); Y factor: both
  R / RX q              xs[q] = zs[q] = {}       (errors before a reset are absorbed)
  CX c t                xs[c] ^= xs[t] ; zs[t] ^= zs[c]
  X_ERROR(p) q          mechanism (p, xs[q])     Z_ERROR(p): (p, zs[q])
  DEPOLARIZE1(p) q      3 independent mechanisms X,Y,Z with q1 = 1/2 - 1/2 sqrt(1 - 4p/3)
  DEPOLARIZE2(p) a b    15 independent mechanisms with q2 = 1/2 - 1/2 (1 - 16p/15)^(1/8)
  M(p)/MX(p)/MPP(p)     mechanism (p, S(m))

Mechanisms with identical symptom sets are merged with p <- p1(1-p2) + p2(1-p1).
Mechanisms with no symptoms are dropped.  Hyper-errors are NOT decomposed.

The output DEM is folded: maximal runs of identical REPEAT iterations (same mechanisms and
detector declarations up to a constant detector shift) become one
`repeat N { ... shift_detectors(dc...) P }` block, which keeps the d25 r100 fixture small and
exercises the repeat / shift_detectors path of the DEM parser.

Usage: circuit_to_dem.py in.stim out.dem [--flat]
"""
import math
import re
import sys


def parse_circuit(text):
    """Returns a flat list of (name, args, targets_str_list, rep_tag) with REPEAT unrolled.

    rep_tag = (block_id, iteration) for instructions inside the outermost REPEAT, else None.
    """
    lines = text.split('\n')
    pos = 0

    def parse_block():
        nonlocal pos
        out = []
        while pos < len(lines):
            line = lines[pos].split('#')[0].strip()
            pos += 1
            if not line:
                continue
            if line == '}':
                return out
            m = re.match(r'^REPEAT\s+(\d+)\s*\{$', line)
            if m:
                body = parse_block()
                out.append(('REPEAT', int(m.group(1)), body))
                continue
            m = re.match(r'^([A-Z_0-9]+)(?:\(([^)]*)\))?\s*(.*)$', line)
            name = m.group(1)
            args = tuple(float(a) for a in m.group(2).split(',')) if m.group(2) else ()
            targets = m.group(3).split()
            out.append((name, args, targets))
        return out

    tree = parse_block()
    flat = []
    block_counter = [0]

    def emit(block, tag):
        for ins in block:
            if ins[0] == 'REPEAT':
                if tag is None:
                    bid = block_counter[0]
                    block_counter[0] += 1
                    for it in range(ins[1]):
                        emit(ins[2], (bid, it))
                else:
                    for it in range(ins[1]):
                        emit(ins[2], tag)
            else:
                flat.append((ins[0], ins[1], ins[2], tag))

    emit(tree, None)
    return flat


def analyse(flat):
    """Returns (num_dets, num_obs, det_coords, det_tag, errors) where errors maps
    symptom tuple -> probability. Symptoms: detector d -> d ; observable k -> -(k+1)."""
    # Forward pass: measurement indices, detectors, observables.
    n_meas = 0
    coord_shift = []
    det_coords = []
    det_tag = []
    det_recs = []
    obs_recs = {}
    meas_count_of = []  # per instruction: number of measurement results it produces
    for name, args, targets, tag in flat:
        k = 0
        if name in ('M', 'MX'):
            k = len(targets)
        elif name == 'MPP':
            k = len(targets)
        elif name == 'DETECTOR':
            recs = [n_meas + int(t[4:-1]) for t in targets]
            c = list(args)
            for i in range(min(len(c), len(coord_shift))):
                c[i] += coord_shift[i]
            det_coords.append(tuple(c))
            det_tag.append(tag)
            det_recs.append(recs)
        elif name == 'OBSERVABLE_INCLUDE':
            recs = [n_meas + int(t[4:-1]) for t in targets]
            obs_recs.setdefault(int(args[0]), []).extend(recs)
        elif name == 'SHIFT_COORDS':
            while len(coord_shift) < len(args):
                coord_shift.append(0.0)
            for i, a in enumerate(args):
                coord_shift[i] += a
        meas_count_of.append(k)
        n_meas += k
    num_dets = len(det_recs)
    num_obs = (max(obs_recs) + 1) if obs_recs else 0

    sens = [set() for _ in range(n_meas)]
    for d, recs in enumerate(det_recs):
        for r in recs:
            sens[r] ^= {d}
    for o, recs in obs_recs.items():
        for r in recs:
            sens[r] ^= {-(o + 1)}
    sens = [frozenset(s) for s in sens]

    xs = {}
    zs = {}
    empty = frozenset()
    errors = {}

    def add_error(p, sym):
        if not sym or p <= 0:
            return
        key = tuple(sorted(sym, key=lambda v: (v < 0, v if v >= 0 else -v)))
        old = errors.get(key)
        if old is None:
            errors[key] = p
        else:
            errors[key] = old * (1 - p) + p * (1 - old)

    m_end = n_meas
    for idx in range(len(flat) - 1, -1, -1):
        name, args, targets, _ = flat[idx]
        k = meas_count_of[idx]
        m_start = m_end - k
        if name == 'CX':
            for i in range(0, len(targets), 2)[::-1]:
                c = int(targets[i])
                t = int(targets[i + 1])
                xt = xs.get(t, empty)
                if xt:
                    xs[c] = xs.get(c, empty) ^ xt
                zc = zs.get(c, empty)
                if zc:
                    zs[t] = zs.get(t, empty) ^ zc
        elif name == 'DEPOLARIZE2':
            p = args[0]
            q = 0.5 - 0.5 * (1 - 16 * p / 15) ** 0.125
            for i in range(0, len(targets), 2)[::-1]:
                a = int(targets[i])
                b = int(targets[i + 1])
                xa = xs.get(a, empty)
                za = zs.get(a, empty)
                xb = xs.get(b, empty)
                zb = zs.get(b, empty)
                pa = (empty, xa, xa ^ za, za)
                pb = (empty, xb, xb ^ zb, zb)
                for ia in range(4):
                    for ib in range(4):
                        if ia or ib:
                            add_error(q, pa[ia] ^ pb[ib])
        elif name == 'DEPOLARIZE1':
            p = args[0]
            q = 0.5 - 0.5 * math.sqrt(1 - 4 * p / 3)
            for t in targets[::-1]:
                a = int(t)
                xa = xs.get(a, empty)
                za = zs.get(a, empty)
                add_error(q, xa)
                add_error(q, xa ^ za)
                add_error(q, za)
        elif name == 'X_ERROR':
            for t in targets[::-1]:
                add_error(args[0], xs.get(int(t), empty))
        elif name == 'Z_ERROR':
            for t in targets[::-1]:
                add_error(args[0], zs.get(int(t), empty))
        elif name in ('M', 'MX'):
            table = xs if name == 'M' else zs
            for i in range(k - 1, -1, -1):
                s = sens[m_start + i]
                if s:
                    qb = int(targets[i])
                    table[qb] = table.get(qb, empty) ^ s
                    if args:
                        add_error(args[0], s)
        elif name == 'MPP':
            for i in range(k - 1, -1, -1):
                s = sens[m_start + i]
                if s:
                    for f in targets[i].split('*'):
                        qb = int(f[1:])
                        if f[0] in 'XY':
                            zs[qb] = zs.get(qb, empty) ^ s
                        if f[0] in 'ZY':
                            xs[qb] = xs.get(qb, empty) ^ s
                    if args:
                        add_error(args[0], s)
        elif name in ('R', 'RX'):
            for t in targets:
                xs.pop(int(t), None)
                zs.pop(int(t), None)
        elif name in ('TICK', 'QUBIT_COORDS', 'DETECTOR', 'OBSERVABLE_INCLUDE', 'SHIFT_COORDS'):
            pass
        else:
            raise ValueError('unsupported instruction ' + name)
        m_end = m_start
    return num_dets, num_obs, det_coords, det_tag, errors


def fmt_num(v):
    if v == int(v):
        return str(int(v))
    return repr(float(v))


def sym_str(sym, off):
    return ' '.join(('D%d' % (s - off)) if s >= 0 else ('L%d' % (-s - 1)) for s in sym)


def write_dem(num_dets, num_obs, det_coords, det_tag, errors, fold=True):
    """Emit DEM text. Errors are grouped by the detector-iteration that holds their smallest
    detector; identical consecutive iterations of one REPEAT block are folded."""
    # Section of each detector: consecutive detectors with equal tag form a section.
    sections = []  # (tag, first_det, end_det)
    for d in range(num_dets):
        if sections and sections[-1][0] == det_tag[d] and det_tag[d] is not None:
            sections[-1][2] = d + 1
        elif sections and det_tag[d] is None and sections[-1][0] is None:
            sections[-1][2] = d + 1
        else:
            sections.append([det_tag[d], d, d + 1])
    sec_of_det = [0] * num_dets
    for si, (_, a, b) in enumerate(sections):
        for d in range(a, b):
            sec_of_det[d] = si
    sec_errors = [[] for _ in sections]
    obs_only = []
    for sym, p in errors.items():
        dets = [s for s in sym if s >= 0]
        if not dets:
            obs_only.append((sym, p))
            continue
        sec_errors[sec_of_det[min(dets)]].append((sym, p))
    for lst in sec_errors:
        lst.sort(key=lambda e: ([s for s in e[0] if s >= 0], [s for s in e[0] if s < 0]))

    def section_signature(si, coord_base):
        _, a, b = sections[si]
        errs = tuple((tuple((s - a) if s >= 0 else s for s in sym), repr(p)) for sym, p in sec_errors[si])
        dets = tuple(tuple(c - cb for c, cb in zip(det_coords[d], coord_base)) for d in range(a, b))
        return (b - a, errs, dets)

    out = []
    id_off = 0
    coord_off = None

    def emit_section(si, indent=''):
        _, a, b = sections[si]
        for sym, p in sec_errors[si]:
            out.append('%serror(%s) %s' % (indent, repr(p), sym_str(sym, id_off)))
        for d in range(a, b):
            c = det_coords[d]
            co = coord_off if coord_off is not None else [0.0] * len(c)
            rel = [c[i] - (co[i] if i < len(co) else 0.0) for i in range(len(c))]
            out.append('%sdetector(%s) D%d' % (indent, ', '.join(fmt_num(v) for v in rel), d - id_off))

    for sym, p in obs_only:
        out.append('error(%s) %s' % (repr(p), sym_str(sym, 0)))
    si = 0
    while si < len(sections):
        tag = sections[si][0]
        run = 1
        if fold and tag is not None and si + 1 < len(sections):
            a0 = sections[si][1]
            base0 = det_coords[a0]
            sig0 = section_signature(si, base0)
            # coordinate delta between consecutive iterations
            while si + run < len(sections) and sections[si + run][0] is not None \
                    and sections[si + run][0][0] == tag[0]:
                a1 = sections[si + run][1]
                if section_signature(si + run, det_coords[a1]) != sig0:
                    break
                run += 1
            if run >= 3:
                a1 = sections[si + 1][1]
                delta = [x - y for x, y in zip(det_coords[a1], det_coords[a0])]
                # All later iterations must share the same delta.
                ok = all(
                    [x - y for x, y in zip(det_coords[sections[si + r + 1][1]], det_coords[sections[si + r][1]])] == delta
                    for r in range(run - 1))
                if not ok:
                    run = 1
            else:
                run = 1
        if run > 1:
            period = sections[si][2] - sections[si][1]
            out.append('repeat %d {' % run)
            emit_section(si, indent='    ')
            out.append('    shift_detectors(%s) %d' % (', '.join(fmt_num(v) for v in delta), period))
            out.append('}')
            id_off += run * period
            if coord_off is None:
                coord_off = [0.0] * len(delta)
            while len(coord_off) < len(delta):
                coord_off.append(0.0)
            for i, v in enumerate(delta):
                coord_off[i] += v * run
            # The DEM parser applies shifts cumulatively; re-anchor so that following ids are relative.
            si += run
        else:
            emit_section(si)
            si += 1
    seen_obs = {-s - 1 for sym in errors for s in sym if s < 0}
    for o in range(num_obs):
        if o not in seen_obs:
            out.append('logical_observable L%d' % o)
    return '\n'.join(out) + '\n'


def circuit_text_to_dem_text(text, fold=True):
    flat = parse_circuit(text)
    nd, no, coords, tags, errors = analyse(flat)
    return write_dem(nd, no, coords, tags, errors, fold=fold)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    fold = '--flat' not in sys.argv
    with open(args[0]) as f:
        text = f.read()
    dem = circuit_text_to_dem_text(text, fold=fold)
    with open(args[1], 'w') as f:
        f.write(dem)


if __name__ == '__main__':
    main()
