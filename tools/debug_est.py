import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import yond_public_b200 as Y
from yond_public_b200.nlf import NlfEstimator
from oracle import yond_oracle as O
rng = np.random.default_rng(11)
imgs = []
for K, S in ((2.0, 3.0), (9.0, 14.0), (5.0, 40.0)):
    imgs.append(np.stack([O.synth_noisy(rng, O.synth_clean_smooth(rng, 128, 128), K, S) for _ in range(32)]))
est = NlfEstimator()
for i, blocks in enumerate(imgs):
    mos = np.concatenate(list(blocks), -1)
    rg = O.bayer2rggb(mos)
    var, mean, lap = O.self_maps(rg, 29)
    th, pct, info = O.get_threshold_score3(lap, mean)
    reg, _ = O._masked_fit(var, mean, lap, th)
    t = Y.bayer2rggb(torch.from_numpy(mos).cuda())[None]
    v2, m2, l2 = est.maps(t, None, 29)
    th2, pct2, info2 = est.threshold_score3(l2, m2)
    reg2, _ = est.masked_fit(v2, m2, l2, th2)
    print(f"img{i}: pct {pct} vs {pct2}; th rel {abs(th-th2)/th:.2e}; reg {reg} vs {reg2}; rel {np.abs(reg-reg2)/np.abs(reg)}")
    print("   map max abs diff var/mean/lap:", float(np.abs(v2[0].cpu().numpy()-var).max()), float(np.abs(m2[0].cpu().numpy()-mean).max()), float(np.abs(l2[0].cpu().numpy()-lap).max()))
    print("   npeaks equal:", np.array_equal(info['npeaks'], info2['npeaks']), "ths max rel", float(np.max(np.abs(info['ths']-info2['ths'])/info['ths'])))
    # fit with oracle maps but our threshold and vice versa
    m = lap < th2
    print("   mask count oracle-th/our-th on oracle maps:", int((lap<th).sum()), int(m.sum()), " on our maps:", int((l2[0].cpu().numpy()<th2).sum()))
