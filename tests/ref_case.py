"""The small hits file shared by the reference-run tests and tools/make_golden_ref_small.py:
150 transcripts, 4000 fragments, three identical-transcript sets (all observed / mixed / all
hit-less members)."""
import os

from mmseq_b200 import hostlib, synth


def make_case(dirname, fmt="text"):
    s = synth.Synth(77, 150, 4000)
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid)
    unobs = [t for t in range(s.T) if h.hdr2col[t] < 0]
    obs = [t for t in range(s.T) if h.hdr2col[t] >= 0]
    ident = [obs[:2], [obs[5], unobs[0], obs[9]], unobs[1:3]]
    path = os.path.join(str(dirname), f"s.{fmt}.hits")
    (synth.write_hits_text if fmt == "text" else synth.write_hits_binary)(s, path, identical=ident)
    return path
