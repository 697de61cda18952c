import importlib, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
import oracle as orc
rng = np.random.default_rng(2026)
bad = 0
shapes = [(int(rng.integers(8, 420)), int(rng.integers(64, 1400))) for _ in range(36)] + [(8, 64), (9, 64), (376, 1248), (377, 1249), (100, 960), (100, 961), (100, 967), (100, 968), (64, 240), (64, 241), (64, 480), (64, 479)]
for (h, w) in shapes:
    levels = 4 if min(h, w) >= 64 else 2
    n = 5
    imgs = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    imgs[1] = 255 * (rng.random((h, w)) > 0.5)
    with pkg.Context(w, h, (500, 500, w / 2, h / 2), levels=levels, max_frames=n, max_jobs=1, max_features=16, max_fa_items=16) as ctx:
        ctx.upload(0, imgs)
        for s in (0, 1, 4):
            ip, gp = orc.build_pyramid(imgs[s], levels)
            ipl, gpl = orc.unpack_pyramid(ip, w, h, levels), orc.unpack_pyramid(gp, w, h, levels)
            for l in range(levels):
                a, b = np.array_equal(ctx.download(s, l, 0), ipl[l]), np.array_equal(ctx.download(s, l, 1), gpl[l])
                if not (a and b):
                    bad += 1
                    print("MISMATCH", (h, w), "frame", s, "level", l, "image ok" if a else "image BAD", "gradient ok" if b else "gradient BAD")
print("shapes", len(shapes), "mismatches", bad)
