#!/usr/bin/env python
"""cuobjdump -sass of the built library -> per-kernel counts of the Blackwell instructions (profiles/rNN_sass_extract.txt).
Usage: cuobjdump -sass semantic_pyramid_for_image_generation_b200/libspyramid_b200.so > /tmp/sass.txt; python tools/sass_extract.py /tmp/sass.txt"""
import collections
import re
import sys

txt = open(sys.argv[1]).read()
parts = re.split(r"\n\s*Function : ", txt)
pat = {'UTCHMMA': r"UTCHMMA(?!\.2CTA)", 'UTCHMMA.2CTA': r"UTCHMMA\.2CTA", 'LDTM': r"\bLDTM", 'UTMALDG': r"UTMALDG",
       'UTMASTG': r"UTMASTG", 'UTCBAR': r"UTCBAR", 'SYNCS': r"SYNCS", 'FPATOM': r"(RED|ATOM)[A-Z.]*\.(F32|F64|F16|BF16)"}
rows = collections.OrderedDict()
fp_total = 0
for p in parts[1:]:
    name = p.split('\n', 1)[0].strip()
    cnt = {k: len(re.findall(v, p)) for k, v in pat.items()}
    fp_total += cnt['FPATOM']
    m = re.search(r"(conv_stack3_kernel|conv_halo2_kernel|conv_halo_kernel|conv_fprop_kernel|conv_wgrad_kernel|"
                  r"wgrad_halo_kernel|sagan_attention_fwd_kernel)", name)
    if not m:
        continue
    k = m.group(1)
    t = re.search(k + r"ILb(\d)ELi(\d+)E", name)
    if t:
        k += "<%s, EPI %s>" % ("split" if t.group(1) == '1' else "bf16", t.group(2))
    rows[k] = cnt
print("cuobjdump -sass semantic_pyramid_for_image_generation_b200/libspyramid_b200.so  (sm_100a, final round-2 build): "
      "instruction counts per kernel")
print("UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA load / "
      "store,\nUTCBAR = tcgen05.commit, SYNCS = mbarrier ops; last column = floating-point global atomics / reductions")
print("(conv_halo / conv_halo2 / conv_stack3 are instantiated per epilogue: EPI 0 = generic, 1..12 = specialised, 13..18 = "
      "pooled;\n one representative of each group is listed, the others differ only in the epilogue's LDG/STG/TMA-store code)")
print("%-44s %8s %13s %6s %8s %8s %7s %6s %s" % ("kernel", "UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR",
                                                "SYNCS", "FP atomics"))
for k, c in rows.items():
    epi = re.search(r"EPI (\d+)", k)
    if epi and int(epi.group(1)) not in (0, 1, 3, 13):
        continue
    print("%-44s %8d %13d %6d %8d %8d %7d %6d %d" % (k, c['UTCHMMA'], c['UTCHMMA.2CTA'], c['LDTM'], c['UTMALDG'],
                                                      c['UTMASTG'], c['UTCBAR'], c['SYNCS'], c['FPATOM']))
print("\nfloating-point atomics anywhere in the library: %d (round 1: split-K, every weight gradient, BN statistics, bias "
      "sums)" % fp_total)
print("kernel functions in the library: %d" % (len(parts) - 1))
