"""Developer tooling: build tile-shape variants of the library side by side (levelsetpy_b200/_hjb200_<tag>.so) so one
GPU call can time them all (tools/time_split.py --lib ...).   python tools/build_variants.py [tag ...]"""
import concurrent.futures as cf
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from levelsetpy_b200 import build as _b  # noqa: E402

P1_OLD = ["HJ_P1_TY=12", "HJ_P1_TXP=21", "HJ_P1_MINB=2", "HJ_P1_GW=0", "HJ_P1_XPAD=0", "HJ_P1_R=8", "HJ_P1_OPT=143"]
P2_OLD = ["HJ_P2_R=6", "HJ_P2_MINB=2", "HJ_P2_VP=8", "HJ_P2_TA=4", "HJ_P2_TB=7", "HJ_P2_GW=0", "HJ_P2_OPT=0"]
FB_OLD = ["HJ_FB_TY=16", "HJ_FB_TXP=16", "HJ_FB_MINB=2", "HJ_FB_GW=0", "HJ_FB_XPAD=0", "HJ_FB_R=8"]


def p1(ty=21, minb=1, gw=2, xpad=8, r=8, opt=143):
    return ["HJ_P1_TY=%d" % ty, "HJ_P1_TXP=21", "HJ_P1_MINB=%d" % minb, "HJ_P1_GW=%d" % gw, "HJ_P1_XPAD=%d" % xpad,
            "HJ_P1_R=%d" % r, "HJ_P1_OPT=%d" % opt]


def p2(r=6, minb=2, vp=8, ta=4, tb=7, gw=1, opt=0):
    return ["HJ_P2_R=%d" % r, "HJ_P2_MINB=%d" % minb, "HJ_P2_VP=%d" % vp, "HJ_P2_TA=%d" % ta, "HJ_P2_TB=%d" % tb,
            "HJ_P2_GW=%d" % gw, "HJ_P2_OPT=%d" % opt]


def fb(ty=8, txp=51, minb=1, gw=1, xpad=0, r=8):
    return ["HJ_FB_TY=%d" % ty, "HJ_FB_TXP=%d" % txp, "HJ_FB_MINB=%d" % minb, "HJ_FB_GW=%d" % gw, "HJ_FB_XPAD=%d" % xpad, "HJ_FB_R=%d" % r]


VARIANTS = {
    "gw4d": ["HJ_P1_4D_GW=1", "HJ_P2_4D_GW=1"],                       # ghost warps in the 4-D pair's kernels (288 threads)
    "p2r7": p2(r=7, ta=4, tb=6),                                      # pass 2: 4 x 6 tile, 7-slot ring (3 planes of slack)
    "p2r8": p2(r=8, ta=4, tb=5),                                      # pass 2: 4 x 5 tile, 8-slot ring (4 planes of slack)
}


def main():
    tags = sys.argv[1:] or list(VARIANTS)
    _b.build()
    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        for so in ex.map(lambda t: _b.build_variant(t, VARIANTS[t]), tags):
            print(so)


if __name__ == "__main__":
    main()
