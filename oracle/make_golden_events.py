"""ORACLE (test infrastructure): mint tests/golden/events_prep.npz by running the UNMODIFIED reference
`Event.queue_raw_events` / `Event.extract_max_refractory_period` / `Event.colorize_events`
(robust_e_nerf/data/datasets.py:132-328) on a synthetic `raw_events.npz`.

    python -m oracle.make_golden_events          # build container only (needs /root/reference)
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import events_prep as EP  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                      "events_prep.npz")


def main():
    ref = ref_loader.load_reference()
    import easydict
    Event = ref.data.datasets.Event
    out = {}
    for tag, (n, w, h, seed, bayer) in {"a": (6000, 24, 16, 1, ""), "b": (4000, 7, 5, 2, "RGGB"),
                                        "c": (50, 64, 48, 3, "")}.items():
        raw = EP.synth_raw_events(n, w, h, seed=seed)
        calib = {"img_height": np.uint16(h), "img_width": np.uint16(w), "bayer_pattern": np.array(bayer)}
        root = tempfile.mkdtemp(prefix="ren_b200_events_")
        np.savez(os.path.join(root, Event.RAW_EVENTS_FILENAME), **raw)
        q = Event.queue_raw_events(root, calib)
        q = Event.colorize_events(easydict.EasyDict(q), calib)
        m = Event.extract_max_refractory_period(raw, calib)
        for k, v in raw.items():
            out[f"{tag}_raw_{k}"] = v
        out[f"{tag}_calib"] = np.array([w, h])
        out[f"{tag}_bayer"] = np.array(bayer)
        for k, v in q.items():
            out[f"{tag}_q_{k}"] = v.numpy()
        out[f"{tag}_max_refractory_period"] = m.numpy()
        print(tag, "events", n, "-> queued", len(q["end_ts"]), "max refractory period", m)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
