"""RGB tile table of the MultiGrid level renderer (`venv.get_images()`, parallel_wrappers.py:187-193 ->
MultiGridEnv.render(mode='level'), multigrid.py:1105-1140 -> Grid.render / render_tile, :159-261).

An adversarial maze only ever holds four kinds of cell -- empty, wall, goal, agent (4 directions) -- each either inside
the agent's highlighted view or not, so a rendered level is a mosaic of 14 distinct 32x32 tiles.  The table is computed
once on the host with numpy (a few hundred microseconds) and a kernel (mgplr_render_images) pastes tiles by index.

Tile = render_tile(obj, highlight=[bool], tile_size=32, subdivs=3): a 96x96 supersampled canvas -- grid lines on the top
and left edges (point_in_rect(0, 0.031, 0, 1) / (0, 1, 0, 0.031), colour 100), the object (wall: all grey 100; goal: all
green (0, 255, 0)), the view highlight for non-wall cells (gym-minigrid highlight_img: img + 0.3 * (uint8(colour) - img)
with the subtraction wrapping in uint8, AGENT_COLOURS[0] = (60, 182, 234)), the agent triangle ((0.12, 0.19), (0.87, 0.50),
(0.12, 0.81) rotated by dir * 90 degrees, drawn after the highlight) -- box-filtered down by 3 and truncated to uint8 when
it is pasted into the image.  gym-minigrid's rendering helpers are third-party (absent here); this is their published
algorithm evaluated for all pixels at once.
"""
import math

import numpy as np

TILE = 32
SUBDIVS = 3
AGENT_COLOUR = np.array([60, 182, 234])   # multigrid.py:46-47
GREY = np.array([100, 100, 100])
GREEN = np.array([0, 255, 0])
N_TILES = 14   # (empty, wall, goal, agent dir 0..3) x (plain, highlighted): index = 2 * code + highlight


def _coords(n):
    c = (np.arange(n) + 0.5) / n
    return np.meshgrid(c, c)   # xf[y, x], yf[y, x]


def _triangle_mask(n, direction):
    """Pixels of Agent.render's triangle (multigrid.py:140-153): point_in_triangle after rotate_fn(theta = dir * pi / 2)."""
    xf, yf = _coords(n)
    theta = 0.5 * math.pi * direction
    cx = cy = 0.5
    x, y = xf - cx, yf - cy
    x2 = cx + x * math.cos(-theta) - y * math.sin(-theta)
    y2 = cy + y * math.cos(-theta) + x * math.sin(-theta)
    a, b, c = np.array((0.12, 0.19)), np.array((0.87, 0.50)), np.array((0.12, 0.81))
    v0, v1 = c - a, b - a
    v2x, v2y = x2 - a[0], y2 - a[1]
    dot00 = v0[0] * v0[0] + v0[1] * v0[1]
    dot01 = v0[0] * v1[0] + v0[1] * v1[1]
    dot02 = v0[0] * v2x + v0[1] * v2y
    dot11 = v1[0] * v1[0] + v1[1] * v1[1]
    dot12 = v1[0] * v2x + v1[1] * v2y
    inv_denom = 1 / (dot00 * dot11 - dot01 * dot01)
    u = (dot11 * dot02 - dot01 * dot12) * inv_denom
    v = (dot00 * dot12 - dot01 * dot02) * inv_denom
    return (u >= 0) & (v >= 0) & ((u + v) < 1)


def render_tile(code, highlight):
    """code: 0 empty, 1 wall, 2 goal, 3 + dir agent.  Returns the uint8 [32, 32, 3] tile as it lands in the image."""
    n = TILE * SUBDIVS
    xf, yf = _coords(n)
    img = np.zeros((n, n, 3), dtype=np.uint8)
    img[(xf >= 0) & (xf <= 0.031)] = (100, 100, 100)
    img[(yf >= 0) & (yf <= 0.031)] = (100, 100, 100)
    if code == 1:
        img[:, :] = GREY
    elif code == 2:
        img[:, :] = GREEN
    if highlight and code != 1:
        blend = img + 0.30 * (np.array(AGENT_COLOUR, dtype=np.uint8) - img)   # uint8 - uint8 wraps, as published
        img[:, :, :] = blend.clip(0, 255).astype(np.uint8)
    if code >= 3:
        img[_triangle_mask(n, code - 3)] = AGENT_COLOUR
    down = img.reshape([TILE, SUBDIVS, TILE, SUBDIVS, 3]).mean(axis=3).mean(axis=1)
    return down.astype(np.uint8)   # `img[ymin:ymax, xmin:xmax, :] = tile_img` truncates the float means


_table = None


def tile_table():
    """uint8 [14, 32, 32, 3]."""
    global _table
    if _table is None:
        _table = np.stack([render_tile(i // 2, bool(i & 1)) for i in range(N_TILES)])
    return _table
