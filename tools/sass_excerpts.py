"""SASS evidence for the tensor-core kernels: python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt
(cuobjdump -sass of the in-tree library; mnemonic counts per kernel and the instruction runs around the first
occurrence of the mnemonics that prove tcgen05 / TMEM / TMA / cluster use)."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pingpong_selfplay_ai_b200", "libpong_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP.S.G.MULTICAST", "UBLKCP", "SYNCS", "UTCBAR.MULTICAST", "UTCBAR", "FFMA2", "FADD2", "UCGABAR",
             "BAR.SYNC", "BAR.RED", "ELECT", "MUFU.EX2", "MUFU.RCP"]
funcs, cur = collections.OrderedDict(), None
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())


def count(lines):
    c = collections.OrderedDict()
    for mn in MNEMONICS:
        k = sum(1 for l in lines if re.search(r"\b" + re.escape(mn) + r"(\b|_)", l.split("*/", 1)[1]))
        if k:
            c[mn] = k
    return c


def excerpt(title, func_pat, mnemonic, before, after, nth=0):
    for name, lines in funcs.items():
        if re.search(func_pat, name):
            hits = [i for i, l in enumerate(lines) if mnemonic in l]
            if len(hits) > nth:
                i = hits[nth]
                print(f"\n## {title}\n# {name}")
                print("\n".join(lines[max(0, i - before): i + after]))
                return
    print(f"\n## {title}\n# (not found: {func_pat} / {mnemonic})")


total = count([l for ls in funcs.values() for l in ls])
print("# SASS evidence for the tensor-core kernels of libpong_b200.so (cuobjdump -sass, sm_100a), round 2")
print("# regenerate: python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt\n")
print("## mnemonic counts per kernel (whole library total: " + ", ".join(f"{k} {v}" for k, v in total.items()) + ")")
print("## (UBLKCP / UTCBAR counts include their .MULTICAST forms)")
for name, lines in funcs.items():
    c = count(lines)
    if any(k in c for k in ("UTCHMMA", "UBLKCP", "UCGABAR", "LDTM")):
        print(name[:110])
        print("    " + ", ".join(f"{k} {v}" for k, v in c.items()))
QN = r"selfplay_tc_kernelIdLb0"
RN = r"selfplay_rnn_tc_kernelIdLb1"
excerpt("fused QNet kernel (f64 env): MMA issue run of one group (L1 pair, then the 13 MMAs of L2: A operand in TMEM -> `tmem[UR..]`, "
        "B descriptors in uniform registers, commit -> mbarrier)", QN, "UTCHMMA", 6, 46)
excerpt("fused QNet kernel: hidden-layer epilogue - tcgen05.ld (LDTM.x16), ReLU + hi/lo split (F2FP.RELU, HADD2.F32, FADD2), "
        "tcgen05.st IN PLACE (STTM.x16)", QN, "STTM", 36, 6)
excerpt("fused QNet kernel: weight blobs by TMA bulk copy (cp.async.bulk -> UBLKCP) on an mbarrier", QN, "UBLKCP", 8, 4)
excerpt("fused QNet kernel: dueling heads as packed fp32 FFMA2 against a broadcast LDS.128 table", QN, "FFMA2", 4, 24, nth=40)
excerpt("fused QNetRNN kernel (paired form): issuer warp - gate MMAs (N = 128) against TMA-streamed weight stages; the slot "
        "release is a commit MULTICAST to both CTAs' `empty` barriers", RN, "UTCBAR.MULTICAST", 30, 3, nth=2)
excerpt("fused QNetRNN kernel (paired form): producer warp - the weight ring; each CTA loads HALF a stage and multicasts it "
        "into both CTAs' rings (cp.async.bulk ... .multicast::cluster)", RN, "UBLKCP.S.G.MULTICAST", 14, 4)
excerpt("fused QNetRNN kernel: LSTM cell on MUFU.EX2 / MUFU.RCP directly (34 instructions, 8 MUFU per unit)", RN, "MUFU.EX2", 4, 40, nth=4)
excerpt("DRQN update, lstm_fwd_kernel: cluster barrier (UCGABAR) after the distributed-shared-memory stores of h_t",
        r"lstm_fwd_kernel", "UCGABAR_ARV", 10, 4, nth=1)
