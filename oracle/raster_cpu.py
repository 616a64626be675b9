"""CPU restatement of the reference's stixel drawing for evaluation (test infrastructure only):
tools/visualization/clustering_visualization.py  draw_stixels :164-414 (label-id image :397-402, disparity result
image :403-409), draw_instance_masks :118-142, instance ids from read_stixel_file :104-111."""
from __future__ import annotations

import numpy as np

# cityscapesscripts/helpers/labels.py: trainId -> id
TRAINID_TO_ID = np.array([7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33], dtype=np.uint8)


def instance_id(semantic_class: int, label: int) -> int:
    """read_stixel_file :104-111 ("convert to cityscapes style"); masks are drawn for ids > 0 only (:123-125)."""
    return semantic_class * 1000 + label if 0 <= label < 1000 else 0


def draw(sections: np.ndarray, instances: dict, rows: int, cols: int, use_cv2: bool = False):
    """sections [C][200] (type == -1 terminates a column), instances {(column, index): label}
    -> (label ids uint8 [rows][cols], instance ids int32, disparity float32)."""
    C = sections.shape[0]
    w = cols // C                                    # stixel_width, :187
    label = np.zeros((rows, cols), np.uint8)
    inst = np.zeros((rows, cols), np.int32)
    disp = np.zeros((rows, cols), np.float32)
    if use_cv2:
        import cv2
    for c in range(C):
        for j in range(sections.shape[1]):
            s = sections[c, j]
            if s["type"] == -1:
                break
            x0, y0 = c * w, rows - int(s["vT"]) - 1  # :205-208, inclusive rectangle
            x1, y1 = x0 + w - 1, rows - int(s["vB"]) - 1
            lid = int(TRAINID_TO_ID[int(s["semantic_class"])])
            iid = instance_id(int(s["semantic_class"]), instances[(c, j)]) if (c, j) in instances else 0
            if use_cv2:
                cv2.rectangle(label, (x0, y0), (x1, y1), lid, thickness=-1)
                cv2.rectangle(inst, (x0, y0), (x1, y1), iid, thickness=-1)
                cv2.rectangle(disp, (x0, y0), (x1, y1), float(s["disparity"]), thickness=-1)
            else:
                label[y0:y1 + 1, x0:x1 + 1] = lid
                inst[y0:y1 + 1, x0:x1 + 1] = iid
                disp[y0:y1 + 1, x0:x1 + 1] = s["disparity"]
    return label, inst, disp
