"""Regenerates the fixtures in tests/golden/ from the REFERENCE tree (run in the build container,
where /root/reference exists; the GPU box only ever reads the committed .npy files).

  rrc_gr3_65_golden.npy   the reference's own known-answer vector for firdes::root_raised_cosine
                          (test/qa_firdes.cpp:10-33: GR3 firdes.root_raised_cosine(1,4,1,0.35,65)),
                          parsed from the reference test source
  rrc_ref_*.npy           outputs of the reference's firdes.hpp itself (compiled in place into
                          oracle/_ref/libref.so) for the parameter sets the hot path uses
  pfb_arb_taps.npy        the reference's 1280 hard-coded resampler taps (PM/pfb_arb_taps.hpp:13),
                          read through oracle/_ref/libref.so
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def main():
    from oracle import pyoracle as po

    po.build(ref=True)
    assert po.ref_lib() is not None, "oracle/_ref/libref.so missing (is /root/reference present?)"
    src = open(os.path.join(REF, "test", "qa_firdes.cpp")).read()
    body = src[src.index("expected_taps = {") + len("expected_taps = {"):]
    body = body[:body.index("};")]
    vals = [float(v) for v in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", body)]
    assert len(vals) == 65
    np.save(os.path.join(HERE, "rrc_gr3_65_golden.npy"), np.array(vals, np.float64))
    np.save(os.path.join(HERE, "rrc_ref_sps4_span11.npy"), po.ref_root_raised_cosine(1.0, 4.0, 1.0, 0.35, 44))
    np.save(os.path.join(HERE, "rrc_ref_gr3_65.npy"), po.ref_root_raised_cosine(1.0, 4.0, 1.0, 0.35, 65))
    np.save(os.path.join(HERE, "rrc_ref_pfb32.npy"), po.ref_root_raised_cosine(32.0, 128.0, 1.0, 0.35, 32 * 4 * 11))
    np.save(os.path.join(HERE, "pfb_arb_taps.npy"), po.ref_pfb_arb_taps())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
