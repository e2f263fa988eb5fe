"""Generate the golden vectors in tests/golden/*.npz.

The reference cannot be imported in this image (scikit-image, rasterio,
geopandas missing), so these vectors are outputs of the CPU
ORACLE (oracle/slic_oracle.py, oracle/stats_oracle.py), not of the reference
itself: they pin the oracle and the CUDA path against regressions, they do not
certify the restatement.  If scikit-image ever becomes importable, regenerate
them with `--skimage` to swap the SLIC call for the real one.

    python tests/golden/make_golden.py [--skimage]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def synth(H, W, C, seed, quantize):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:H, :W].astype(np.float64)
    bands = []
    for c in range(C):
        fy, fx, ph = rng.uniform(0.02, 0.1, 3)
        bands.append(np.sin(yy * fy + ph) + np.cos(xx * fx - ph) + 0.5 * np.sin((yy - xx) * fy))
    img = np.stack(bands, -1)
    img = (img - img.min()) / (img.max() - img.min()) + rng.normal(0, 0.04, img.shape)
    if quantize:
        return np.clip(np.round(img * 255), 0, 255).astype(np.uint8)
    return img.astype(np.float32)


def main():
    import slic_oracle as so
    import stats_oracle
    use_skimage = "--skimage" in sys.argv
    if use_skimage:
        from skimage.segmentation import slic as sk_slic

    def labels_of(raw, bands, **kw):
        if not use_skimage:
            return so.create_segments_labels(raw.astype(np.float32), bands, **kw)
        img = raw.astype(np.float32)
        for i in range(img.shape[2]):
            img[:, :, i] = so.normalize_band(img[:, :, i])
        seg = sk_slic(img[:, :, bands], **kw)
        if kw.get("mask") is not None:
            seg[kw["mask"] == 0] = -1
        return seg

    cases = {
        "slic_rgb_64": dict(raw=synth(64, 64, 3, 1, True), bands=[0, 1, 2],
                            kw=dict(n_segments=30, compactness=10.0)),
        "slic_ms8_96": dict(raw=synth(96, 80, 8, 2, False), bands=[0, 2, 3, 5, 7],
                            kw=dict(n_segments=50, compactness=0.1, max_num_iter=10)),
        "slic_masked_80": dict(raw=synth(80, 80, 4, 3, False), bands=[0, 1, 2, 3],
                               kw=dict(n_segments=20, compactness=0.3), masked=True),
    }
    for name, c in cases.items():
        raw, bands, kw = c["raw"], c["bands"], dict(c["kw"])
        save = {}
        if c.get("masked"):
            yy, xx = np.mgrid[:raw.shape[0], :raw.shape[1]]
            mask = ((yy - 40) ** 2 + (xx - 38) ** 2) < 33 ** 2
            kw["mask"] = mask
            save["mask"] = mask.astype(np.uint8)
        labels = labels_of(raw, bands, **kw)
        ids = np.unique(labels[labels >= 0])
        stats, counts = stats_oracle.zonal_stats(labels, raw.astype(np.float32), list(range(raw.shape[2])), ids)
        for k, v in c["kw"].items():
            save["kw_" + k] = np.asarray(v)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), raw=raw, bands=np.asarray(bands),
                            labels=labels.astype(np.int32), ids=ids, stats=stats, counts=counts, **save)
        print(name, raw.shape, raw.dtype, "labels", len(ids))
    make_texture_golden()


def make_texture_golden():
    """GLCM texture features (oracle/texture_oracle.py) of the segments of two fixtures: a float32
    raster (float32 quantisation) and the uint8 raster (float64 quantisation, like the reference's
    masked crop of an integer raster)."""
    import texture_oracle
    for src, bands in (("slic_ms8_96", [0, 3, 7]), ("slic_rgb_64", [0, 1, 2])):
        z = np.load(os.path.join(HERE, src + ".npz"))
        raw, labels, ids = z["raw"], z["labels"], z["ids"]
        f64 = raw.dtype != np.float32
        tex = texture_oracle.textural_stats(labels, raw.astype(np.float32), bands, ids,
                                            compute_dtype=np.float64 if f64 else np.float32)
        np.savez_compressed(os.path.join(HERE, "texture_" + src[5:] + ".npz"), source=np.asarray(src),
                            bands=np.asarray(bands), ids=ids, texture=tex, quantise_f64=np.asarray(f64))
        print("texture", src, tex.shape)


if __name__ == "__main__":
    main()
