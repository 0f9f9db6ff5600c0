import numpy as np

_NAMED = {"green": (0.0, 0.5, 0.0), "yellow": (1.0, 1.0, 0.0), "red": (1.0, 0.0, 0.0), "blue": (0.0, 0.0, 1.0),
          "white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0)}


class LinearSegmentedColormap:
    def __init__(self, name, stops):
        self.name = name
        self.pos = np.array([p for p, _ in stops], dtype=np.float64)
        self.rgb = np.array([_NAMED[c] if isinstance(c, str) else c for _, c in stops], dtype=np.float64)

    @classmethod
    def from_list(cls, name, colors, N=256):
        if not isinstance(colors[0], (tuple, list)) or isinstance(colors[0][0], str):
            colors = [(i / (len(colors) - 1), c) for i, c in enumerate(colors)]
        return cls(name, colors)

    def __call__(self, x):
        x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
        out = np.stack([np.interp(x, self.pos, self.rgb[:, c]) for c in range(3)] + [np.ones_like(x)], -1)
        return out
