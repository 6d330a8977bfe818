"""Regenerate tests/golden/* from the read-only reference checkout (run in the build container;
/root/reference does not exist on the GPU box, which is why the results are committed).

  kitti_<pair>_{10,11}_gray.png   the two bundled frame pairs the reference author ran
                                  (HornSchunckOF/main.cpp:42-43 and the 000040 twin), after the
                                  reference's own preprocess() (main.cpp:11-26, BGR2GRAY)
  plot_<pair>.npz                 the reference's golden artefact
                                  HornSchunckOF/img/resimage/<pair>_10.pnghsbresenhamLineFlow.png
                                  reduced to the pixels plotFlow.cpp drew (positions + colours),
                                  i.e. golden != imagePrevRaw; `base_at` holds the raw colours
                                  at those positions so the comparison needs no colour frame
"""
import os
import sys

import cv2
import numpy as np

REF = "/root/reference/HornSchunckOF/img"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    for pair in ("000040", "000050"):
        raw_prev = cv2.imread(f"{REF}/leftimage/{pair}_10.png")
        raw_next = cv2.imread(f"{REF}/leftimage/{pair}_11.png")
        saved_prev = cv2.imread(f"{REF}/resimage/{pair}_10.pngimagePrevRaw.png")
        saved_next = cv2.imread(f"{REF}/resimage/{pair}_10.pngimageNextRaw.png")
        assert np.array_equal(raw_prev, saved_prev) and np.array_equal(raw_next, saved_next)
        for tag, img in (("10", raw_prev), ("11", raw_next)):
            gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
            cv2.imwrite(f"{OUT}/kitti_{pair}_{tag}_gray.png", gray, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        gold = cv2.imread(f"{REF}/resimage/{pair}_10.pnghsbresenhamLineFlow.png")
        diff = (gold != raw_prev).any(axis=2)
        yx = np.argwhere(diff).astype(np.int16)
        np.savez_compressed(f"{OUT}/plot_{pair}.npz", yx=yx, bgr=gold[diff], base_at=raw_prev[diff],
                            shape=np.array(gold.shape[:2], np.int32))
        print(pair, "drawn pixels:", len(yx))


if __name__ == "__main__":
    sys.exit(main())
