"""Build tuning variants of libmqi_b200.so (block size / resident CTAs / loop barrier) into
moquimc_b200/variants/ for A/B runs on the GPU box:  MQI_B200_LIB=<variant.so> python scripts/quick_bench.py"""
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moquimc_b200 import build as B

VARIANTS = {
    "m640": (768, 1, 0, "-DMQI_K_BLOCK_MULTI=640"), "m512": (768, 1, 0, "-DMQI_K_BLOCK_MULTI=512"), "m768": (768, 1, 0),
    "rspx": (768, 1, 0, "-DMQI_K_RSP_EXACT=1"), "pw3": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=3"),
    "adv1": (768, 1, 0, "-DMQI_K_ADV_BATCH=1", "-DMQI_K_ADV_TURNS=0"), "adv10": (768, 1, 0, "-DMQI_K_ADV_BATCH=6", "-DMQI_K_ADV_TURNS=10", "-DMQI_K_BLOCK_MULTI=640"),
    "pw1": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=1"), "pw2": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=2"),
    "pw4": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=4"), "adv10b": (768, 1, 0, "-DMQI_K_ADV_BATCH=6", "-DMQI_K_ADV_TURNS=10"),
        "cur640": (768, 1, 0, "-DMQI_K_BLOCK_DIJ=640"),
    "old": (768, 1, 0, "-DMQI_K_ADV_QUEUE=0", "-DMQI_K_DIJ_COOP=0"), "new": (768, 1, 0),
    "coop_b2": (768, 1, 0, "-DMQI_K_DIJ_BATCHES=2"), "coop640": (768, 1, 0, "-DMQI_K_BLOCK_DIJ=640"),
    "coop640_b2": (768, 1, 0, "-DMQI_K_BLOCK_DIJ=640", "-DMQI_K_DIJ_BATCHES=2"), "coop_t4b2": (768, 1, 0, "-DMQI_K_DIJ_TEAM=4", "-DMQI_K_DIJ_BATCHES=2"),
    "coop_t4b3": (768, 1, 0, "-DMQI_K_DIJ_TEAM=4", "-DMQI_K_DIJ_BATCHES=3"), "coop_t16": (768, 1, 0, "-DMQI_K_DIJ_TEAM=16", "-DMQI_K_DIJ_BATCHES=3"),
    "advmin16": (768, 1, 0, "-DMQI_K_ADV_MIN=16"), "advmin24": (768, 1, 0, "-DMQI_K_ADV_MIN=24"), "adv768": (768, 1, 0, "-DMQI_K_BLOCK_MULTI=768"),
    "pw2_4": (768, 1, 0, "-DMQI_K_PROBE_WIDTH2=4"), "pw2_6": (768, 1, 0, "-DMQI_K_PROBE_WIDTH2=6"), "pw2_8": (768, 1, 0, "-DMQI_K_PROBE_WIDTH2=8"),
    "pw1_4": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=1", "-DMQI_K_PROBE_WIDTH2=4"), "pw3_6": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=3", "-DMQI_K_PROBE_WIDTH2=6"),
    "pw1_2": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=1", "-DMQI_K_PROBE_WIDTH2=2"), "pw1_3": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=1", "-DMQI_K_PROBE_WIDTH2=3"),
    "pw1_5": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=1", "-DMQI_K_PROBE_WIDTH2=5"), "pw1_6": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=1", "-DMQI_K_PROBE_WIDTH2=6"),
    "pw1_8": (768, 1, 0, "-DMQI_K_PROBE_WIDTH=1", "-DMQI_K_PROBE_WIDTH2=8"),
    "pw2_3": (768, 1, 0, "-DMQI_K_PROBE_WIDTH2=3"),
    "m576": (768, 1, 0, "-DMQI_K_BLOCK_MULTI=576"), "m704": (768, 1, 0, "-DMQI_K_BLOCK_MULTI=704"),
    "apxdiv": (768, 1, 0, "-DMQI_K_EXACT_DIV=0"),
    "cur896": (768, 1, 0, "-DMQI_K_BLOCK_DIJ=896"), "cur1024": (768, 1, 0, "-DMQI_K_BLOCK_DIJ=1024"),
    "cur": (768, 1, 0), "ph10": (768, 1, 0, "-DMQI_K_PHILOX_ROUNDS=10"),
    "b256": (256, 4, 0), "b512": (512, 2, 0), "b128": (128, 8, 0), "b256_3": (256, 3, 0),
    "early_lut": (256, 4, 0, "-DMQI_K_LATE_LUT=0"), "base": (256, 4, 0),
    "b256_4": (256, 4, 0), "early_lut3": (256, 3, 0, "-DMQI_K_LATE_LUT=0"), "b512_1": (512, 1, 0), "b384_2": (384, 2, 0), "b768_1": (768, 1, 0), "dlcm_cg": (768, 1, 0, "-Xptxas", "-dlcm=cg"), "dlcm_ca": (768, 1, 0, "-Xptxas", "-dlcm=ca"),
    "b640_1": (640, 1, 0), "b896_1": (896, 1, 0), "b1024_1": (1024, 1, 0), "b832_1": (832, 1, 0), "b448_2": (448, 2, 0),
    "b128_5": (128, 5, 0), "b320_2": (320, 2, 0), "b224_3": (224, 3, 0), "b192_3": (192, 3, 0), "b352_2": (352, 2, 0),
}

def main(names):
    B.gen_tables_inc()
    out = os.path.join(B.HERE, "variants")
    os.makedirs(out, exist_ok=True)
    cu = [os.path.join(B.CSRC, f) for f in ("mqi_transport.cu", "mqi_capi.cu")]
    procs = []
    for n in names:
        blk, mb, sync = VARIANTS[n][:3]
        cmd = [B.nvcc()] + B.NVCC_FLAGS + list(VARIANTS[n][3:]) + ["-DMQI_K_BLOCK=%d" % blk, "-DMQI_K_MIN_BLOCKS=%d" % mb, 
                                           "-shared", "-o", os.path.join(out, "libmqi_%s.so" % n)] + cu + ["-ldl"]
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        assert p.wait() == 0

if __name__ == "__main__":
    main(sys.argv[1:] or list(VARIANTS))
