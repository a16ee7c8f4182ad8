import numpy as np, sys
sys.path.insert(0,'/root/repo')
from groundsegmentationrosnode_b200 import capi, params as P, synth
from oracle import oracle as O
for preset in (P.GS_PRESET_VLP16, P.GS_PRESET_HDL64E_2048, P.GS_PRESET_OS1_128, P.GS_PRESET_CLOUD2RANGE_DEFAULT):
    p = P.preset(preset)
    pts = synth.make_adversarial_frame(p, synth.SEED0 + 3, "collisions") if preset == P.GS_PRESET_VLP16 else synth.make_frame(p, synth.SEED0, shuffle=(preset==P.GS_PRESET_OS1_128))
    want = O.process_frame(p, pts)
    with capi.Context(p) as ctx:
        got = ctx.process_frame(pts)
        ok = all(np.array_equal(got[k], want[k]) for k in ("range","angle","no_ground","xyz"))
        print(preset, pts.shape[0], "ok" if ok else "MISMATCH", ctx.stats().kernel_launches)
