"""Golden vectors for the video-level helpers, produced by running the REFERENCE's own
codes/data/util.py::index_generation (data/util.py:169-214) in this container.
    python tests/golden/make_golden_video.py      -> tests/golden/index_generation.json"""
import json
import os
import sys

REF = os.environ.get("RVSR_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "codes"))
import data.util as ref_util  # noqa: E402  (needs cv2 / PIL, present in the build container)

cases = []
for padding in ("replicate", "reflection", "new_info", "circle"):
    for N in (3, 5, 7):
        for max_n in (7, 10, 50):
            for crt in sorted(set([0, 1, 2, 3, max_n // 2, max_n - 4, max_n - 3, max_n - 2, max_n - 1])):
                if 0 <= crt < max_n:
                    cases.append(dict(crt=crt, max_n=max_n, N=N, padding=padding,
                                      out=ref_util.index_generation(crt, max_n, N, padding=padding)))
json.dump(cases, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "index_generation.json"), "w"))
print(len(cases), "cases")
