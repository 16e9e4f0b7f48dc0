"""No-op stand-ins for gym_minigrid.rendering (RGB tile rendering is out of scope)."""


def _noop(*a, **k):
    return None


point_in_triangle = rotate_fn = fill_coords = point_in_rect = _noop
highlight_img = point_in_circle = point_in_line = _noop


def downsample(img, factor):
    return img
