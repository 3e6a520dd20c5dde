"""Tiles-per-group sweep of the tcgen05 convolution at the bench's level shapes (class balance: T tiles are dealt to
NM = 1, 2 or 4 issuing classes, so T = 5 or 3 leaves one class with twice the work of the others).
    python tools/tc_tsweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0], "quick", "99"]            # import tc_sweep's helpers without running its sweep
import tc_sweep as S

for shp, ts in (((154605, 27, 96, 96), (0, 2, 3, 4, 5)), ((138398, 27, 96, 96), (0, 4, 5)), ((20727, 27, 160, 160), (0, 1, 2, 3)),
                ((18673, 27, 160, 160), (0, 2, 3)), ((59700, 27, 128, 128), (0, 2, 3, 4)), ((317485, 27, 64, 64), (0, 4, 6, 8)),
                ((154605, 8, 64, 96), (0, 2, 4, 5)), ((59700, 8, 96, 128), (0, 2, 4)), ((154605, 8, 128, 96), (0, 2, 4, 5))):
    c = S.make(*shp)
    for t in ts:
        S.run(c, t=t)
    del c
    S.torch.cuda.empty_cache()
