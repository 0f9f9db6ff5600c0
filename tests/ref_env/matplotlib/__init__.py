"""matplotlib stand-in for the reference's heat-map writers (scripts/map_generator.py:22-66): colormap evaluation and figure
saving through numpy + PIL; plotting is outside the hot path."""


def use(*a, **kw):
    pass
