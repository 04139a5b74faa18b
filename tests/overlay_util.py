"""Shared by the overlay tests: random face lists that stress the clipping / ordering rules, the host drawing they are
compared with, and a numpy executor of a DrawList (checks the lowering on the CPU box, without the kernel)."""
import numpy as np


def random_faces(rng, h, w, n_frames, max_faces=6):
    """[(bbox int32[4], name | "Unknown", similarity)] per frame: boxes inside, across and outside the frame borders,
    tiny and inverted boxes, overlapping faces."""
    names = ["Ann", "Bob", "Chandrasekhar", "D"]
    out = []
    for _ in range(n_frames):
        faces = []
        for _ in range(int(rng.integers(0, max_faces + 1))):
            kind = int(rng.integers(0, 6))
            if kind == 0:      # ordinary box
                x1, y1 = int(rng.integers(0, w - 40)), int(rng.integers(0, h - 40))
                x2, y2 = x1 + int(rng.integers(8, 200)), y1 + int(rng.integers(8, 200))
            elif kind == 1:    # crosses a border
                x1, y1 = int(rng.integers(-120, 30)), int(rng.integers(-120, 30))
                x2, y2 = x1 + int(rng.integers(20, 200)), y1 + int(rng.integers(20, 200))
            elif kind == 2:    # far corner
                x2, y2 = int(rng.integers(w - 30, w + 150)), int(rng.integers(h - 30, h + 150))
                x1, y1 = x2 - int(rng.integers(20, 200)), y2 - int(rng.integers(20, 200))
            elif kind == 3:    # tiny (corner length 0 or 1)
                x1, y1 = int(rng.integers(0, w)), int(rng.integers(0, h))
                x2, y2 = x1 + int(rng.integers(0, 9)), y1 + int(rng.integers(0, 9))
            elif kind == 4:    # inverted
                x1, y1 = int(rng.integers(50, w)), int(rng.integers(50, h))
                x2, y2 = x1 - int(rng.integers(1, 60)), y1 - int(rng.integers(1, 60))
            else:              # wholly outside
                x1, y1 = w + int(rng.integers(5, 500)), -int(rng.integers(300, 900))
                x2, y2 = x1 + 50, y1 + 50
            bbox = np.array([x1, y1, x2, y2], np.int32)
            if rng.random() < 0.6:
                faces.append((bbox, names[int(rng.integers(0, len(names)))], float(np.float32(rng.uniform(0.4, 1.0)))))
            else:
                faces.append((bbox, "Unknown", 0.0))
        out.append(faces)
    return out


COLORS = {"Ann": (10, 200, 30), "Bob": (256, 0, 128), "Chandrasekhar": (0, 0, 255), "D": (77, 77, 77)}


def draw_host(helpers, frames, faces, colors=COLORS):
    """The reference loop's drawing (main.py:144-148) with `helpers` = the reference's or this repo's utils.helpers."""
    for frame, per_frame in zip(frames, faces):
        for bbox, name, sim in per_frame:
            if name != "Unknown":
                helpers.draw_bbox_info(frame, bbox, similarity=sim, name=name, color=colors[name])
            else:
                helpers.draw_bbox(frame, bbox, (255, 0, 0))
    return frames


def execute_draw_list(dl, frames):
    """numpy execution of the packed draw list, command by command, in order."""
    cmds, frame_groups, group_cmds, masks = dl.pack()
    h, w = frames[0].shape[:2]
    for f, frame in enumerate(frames):
        for g in range(frame_groups[f], frame_groups[f + 1]):
            for kind, x0, y0, x1, y1, bgr, off, _ in cmds[group_cmds[g]:group_cmds[g + 1]]:
                color = (bgr & 255, (bgr >> 8) & 255, (bgr >> 16) & 255)
                if kind == 0:
                    x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, w - 1), min(y1, h - 1)
                    if x0 <= x1 and y0 <= y1:
                        frame[y0:y1 + 1, x0:x1 + 1] = color
                else:
                    m = masks[off:off + x1 * y1].reshape(y1, x1).astype(bool)
                    ys, xs = np.nonzero(m)
                    ys, xs = ys + y0, xs + x0
                    ok = (ys >= 0) & (ys < h) & (xs >= 0) & (xs < w)
                    frame[ys[ok], xs[ok]] = color
    return frames
