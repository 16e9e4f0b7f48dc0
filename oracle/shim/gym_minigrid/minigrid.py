"""gym-minigrid 1.0.1 `minigrid.py`, restated (see package docstring)."""
import math
from enum import IntEnum

import numpy as np

import gym
from gym import spaces
from gym.utils import seeding

TILE_PIXELS = 32

COLORS = {
    'red': np.array([255, 0, 0]),
    'green': np.array([0, 255, 0]),
    'blue': np.array([0, 0, 255]),
    'purple': np.array([112, 39, 195]),
    'yellow': np.array([255, 255, 0]),
    'grey': np.array([100, 100, 100]),
}
COLOR_NAMES = sorted(list(COLORS.keys()))
COLOR_TO_IDX = {'red': 0, 'green': 1, 'blue': 2, 'purple': 3, 'yellow': 4, 'grey': 5}
IDX_TO_COLOR = dict(zip(COLOR_TO_IDX.values(), COLOR_TO_IDX.keys()))
OBJECT_TO_IDX = {
    'unseen': 0, 'empty': 1, 'wall': 2, 'floor': 3, 'door': 4, 'key': 5,
    'ball': 6, 'box': 7, 'goal': 8, 'lava': 9, 'agent': 10,
}
IDX_TO_OBJECT = dict(zip(OBJECT_TO_IDX.values(), OBJECT_TO_IDX.keys()))
STATE_TO_IDX = {'open': 0, 'closed': 1, 'locked': 2}
DIR_TO_VEC = [
    np.array((1, 0)),   # right (+x)
    np.array((0, 1)),   # down (+y)
    np.array((-1, 0)),  # left
    np.array((0, -1)),  # up
]


class WorldObj(object):
    def __init__(self, type, color):
        assert type in OBJECT_TO_IDX, type
        assert color in COLOR_TO_IDX, color
        self.type = type
        self.color = color
        self.contains = None
        self.init_pos = None
        self.cur_pos = None

    def can_overlap(self):
        return False

    def can_pickup(self):
        return False

    def can_contain(self):
        return False

    def see_behind(self):
        return True

    def toggle(self, env, pos):
        return False

    def encode(self):
        return (OBJECT_TO_IDX[self.type], COLOR_TO_IDX[self.color], 0)

    def render(self, r):
        pass


class Goal(WorldObj):
    def __init__(self):
        super().__init__('goal', 'green')

    def can_overlap(self):
        return True

    def render(self, img):
        from gym_minigrid.rendering import fill_coords, point_in_rect
        fill_coords(img, point_in_rect(0, 1, 0, 1), COLORS[self.color])


class Floor(WorldObj):
    def __init__(self, color='blue'):
        super().__init__('floor', color)

    def can_overlap(self):
        return True


class Lava(WorldObj):
    def __init__(self):
        super().__init__('lava', 'red')

    def can_overlap(self):
        return True


class Wall(WorldObj):
    def __init__(self, color='grey'):
        super().__init__('wall', color)

    def see_behind(self):
        return False

    def render(self, img):
        from gym_minigrid.rendering import fill_coords, point_in_rect
        fill_coords(img, point_in_rect(0, 1, 0, 1), COLORS[self.color])


class Door(WorldObj):
    def __init__(self, color, is_open=False, is_locked=False):
        super().__init__('door', color)
        self.is_open = is_open
        self.is_locked = is_locked

    def can_overlap(self):
        return self.is_open

    def see_behind(self):
        return self.is_open

    def toggle(self, env, pos):
        if self.is_locked:
            if isinstance(env.carrying, Key) and env.carrying.color == self.color:
                self.is_locked = False
                self.is_open = True
                return True
            return False
        self.is_open = not self.is_open
        return True

    def encode(self):
        if self.is_open:
            state = 0
        elif self.is_locked:
            state = 2
        else:
            state = 1
        return (OBJECT_TO_IDX[self.type], COLOR_TO_IDX[self.color], state)


class Key(WorldObj):
    def __init__(self, color='blue'):
        super().__init__('key', color)

    def can_pickup(self):
        return True


class Ball(WorldObj):
    def __init__(self, color='blue'):
        super().__init__('ball', color)

    def can_pickup(self):
        return True


class Box(WorldObj):
    def __init__(self, color, contains=None):
        super().__init__('box', color)
        self.contains = contains

    def can_pickup(self):
        return True

    def toggle(self, env, pos):
        env.grid.set(*pos, self.contains)
        return True


class Grid(object):
    tile_cache = {}

    def __init__(self, width, height):
        assert width >= 3
        assert height >= 3
        self.width = width
        self.height = height
        self.grid = [None] * width * height

    def set(self, i, j, v):
        assert i >= 0 and i < self.width
        assert j >= 0 and j < self.height
        self.grid[j * self.width + i] = v

    def get(self, i, j):
        assert i >= 0 and i < self.width
        assert j >= 0 and j < self.height
        return self.grid[j * self.width + i]

    def horz_wall(self, x, y, length=None, obj_type=Wall):
        if length is None:
            length = self.width - x
        for i in range(0, length):
            self.set(x + i, y, obj_type())

    def vert_wall(self, x, y, length=None, obj_type=Wall):
        if length is None:
            length = self.height - y
        for j in range(0, length):
            self.set(x, y + j, obj_type())

    def wall_rect(self, x, y, w, h):
        self.horz_wall(x, y, w)
        self.horz_wall(x, y + h - 1, w)
        self.vert_wall(x, y, h)
        self.vert_wall(x + w - 1, y, h)

    def rotate_left(self):
        grid = Grid(self.height, self.width)
        for i in range(self.width):
            for j in range(self.height):
                v = self.get(i, j)
                grid.set(j, grid.height - 1 - i, v)
        return grid

    def slice(self, topX, topY, width, height):
        grid = Grid(width, height)
        for j in range(0, height):
            for i in range(0, width):
                x = topX + i
                y = topY + j
                if x >= 0 and x < self.width and y >= 0 and y < self.height:
                    v = self.get(x, y)
                else:
                    v = Wall()
                grid.set(i, j, v)
        return grid

    def encode(self, vis_mask=None):
        if vis_mask is None:
            vis_mask = np.ones((self.width, self.height), dtype=bool)
        array = np.zeros((self.width, self.height, 3), dtype='uint8')
        for i in range(self.width):
            for j in range(self.height):
                if vis_mask[i, j]:
                    v = self.get(i, j)
                    if v is None:
                        array[i, j, 0] = OBJECT_TO_IDX['empty']
                        array[i, j, 1] = 0
                        array[i, j, 2] = 0
                    else:
                        array[i, j, :] = v.encode()
        return array

    def process_vis(grid, agent_pos):
        mask = np.zeros(shape=(grid.width, grid.height), dtype=bool)
        mask[agent_pos[0], agent_pos[1]] = True
        for j in reversed(range(0, grid.height)):
            for i in range(0, grid.width - 1):
                if not mask[i, j]:
                    continue
                cell = grid.get(i, j)
                if cell and not cell.see_behind():
                    continue
                mask[i + 1, j] = True
                if j > 0:
                    mask[i + 1, j - 1] = True
                    mask[i, j - 1] = True
            for i in reversed(range(1, grid.width)):
                if not mask[i, j]:
                    continue
                cell = grid.get(i, j)
                if cell and not cell.see_behind():
                    continue
                mask[i - 1, j] = True
                if j > 0:
                    mask[i - 1, j - 1] = True
                    mask[i, j - 1] = True
        for j in range(0, grid.height):
            for i in range(0, grid.width):
                if not mask[i, j]:
                    grid.set(i, j, None)
        return mask


class MiniGridEnv(gym.Env):
    """Single-agent MiniGrid base class (gym-minigrid 1.0.1), restated."""

    metadata = {'render.modes': ['human', 'rgb_array'], 'video.frames_per_second': 10}

    class Actions(IntEnum):
        left = 0
        right = 1
        forward = 2
        pickup = 3
        drop = 4
        toggle = 5
        done = 6

    def __init__(self, grid_size=None, width=None, height=None, max_steps=100,
                 see_through_walls=False, seed=1337, agent_view_size=7):
        if grid_size:
            assert width is None and height is None
            width = grid_size
            height = grid_size
        self.actions = MiniGridEnv.Actions
        self.action_space = spaces.Discrete(len(self.actions))
        self.agent_view_size = agent_view_size
        self.observation_space = spaces.Box(
            low=0, high=255, shape=(self.agent_view_size, self.agent_view_size, 3), dtype='uint8')
        self.observation_space = spaces.Dict({'image': self.observation_space})
        self.reward_range = (0, 1)
        self.window = None
        self.width = width
        self.height = height
        self.max_steps = max_steps
        self.see_through_walls = see_through_walls
        self.agent_pos = None
        self.agent_dir = None
        self.seed(seed=seed)
        self.reset()

    def reset(self):
        self.agent_pos = None
        self.agent_dir = None
        self._gen_grid(self.width, self.height)
        assert self.agent_pos is not None
        assert self.agent_dir is not None
        start_cell = self.grid.get(*self.agent_pos)
        assert start_cell is None or start_cell.can_overlap()
        self.carrying = None
        self.step_count = 0
        obs = self.gen_obs()
        return obs

    def seed(self, seed=1337):
        self.np_random, _ = seeding.np_random(seed)
        return [seed]

    @property
    def steps_remaining(self):
        return self.max_steps - self.step_count

    def _gen_grid(self, width, height):
        assert False, "_gen_grid needs to be implemented by each environment"

    def _reward(self):
        return 1 - 0.9 * (self.step_count / self.max_steps)

    def _rand_int(self, low, high):
        return self.np_random.randint(low, high)

    def _rand_float(self, low, high):
        return self.np_random.uniform(low, high)

    def _rand_bool(self):
        return (self.np_random.randint(0, 2) == 0)

    def _rand_elem(self, iterable):
        lst = list(iterable)
        idx = self._rand_int(0, len(lst))
        return lst[idx]

    def _rand_subset(self, iterable, num_elems):
        lst = list(iterable)
        assert num_elems <= len(lst)
        out = []
        while len(out) < num_elems:
            elem = self._rand_elem(lst)
            lst.remove(elem)
            out.append(elem)
        return out

    def _rand_color(self):
        return self._rand_elem(COLOR_NAMES)

    def _rand_pos(self, xLow, xHigh, yLow, yHigh):
        return (self.np_random.randint(xLow, xHigh), self.np_random.randint(yLow, yHigh))

    def place_obj(self, obj, top=None, size=None, reject_fn=None, max_tries=math.inf):
        if top is None:
            top = (0, 0)
        else:
            top = (max(top[0], 0), max(top[1], 0))
        if size is None:
            size = (self.grid.width, self.grid.height)
        num_tries = 0
        while True:
            if num_tries > max_tries:
                raise RecursionError('rejection sampling failed in place_obj')
            num_tries += 1
            pos = np.array((
                self._rand_int(top[0], min(top[0] + size[0], self.grid.width)),
                self._rand_int(top[1], min(top[1] + size[1], self.grid.height))))
            if self.grid.get(*pos) is not None:
                continue
            if np.array_equal(pos, self.agent_pos):
                continue
            if reject_fn and reject_fn(self, pos):
                continue
            break
        self.grid.set(*pos, obj)
        if obj is not None:
            obj.init_pos = pos
            obj.cur_pos = pos
        return pos

    def put_obj(self, obj, i, j):
        self.grid.set(i, j, obj)
        obj.init_pos = (i, j)
        obj.cur_pos = (i, j)

    def place_agent(self, top=None, size=None, rand_dir=True, max_tries=math.inf):
        self.agent_pos = None
        pos = self.place_obj(None, top, size, max_tries=max_tries)
        self.agent_pos = pos
        if rand_dir:
            self.agent_dir = self._rand_int(0, 4)
        return pos

    @property
    def dir_vec(self):
        assert self.agent_dir >= 0 and self.agent_dir < 4
        return DIR_TO_VEC[self.agent_dir]

    @property
    def right_vec(self):
        dx, dy = self.dir_vec
        return np.array((-dy, dx))

    @property
    def front_pos(self):
        return self.agent_pos + self.dir_vec

    def get_view_exts(self):
        if self.agent_dir == 0:
            topX = self.agent_pos[0]
            topY = self.agent_pos[1] - self.agent_view_size // 2
        elif self.agent_dir == 1:
            topX = self.agent_pos[0] - self.agent_view_size // 2
            topY = self.agent_pos[1]
        elif self.agent_dir == 2:
            topX = self.agent_pos[0] - self.agent_view_size + 1
            topY = self.agent_pos[1] - self.agent_view_size // 2
        elif self.agent_dir == 3:
            topX = self.agent_pos[0] - self.agent_view_size // 2
            topY = self.agent_pos[1] - self.agent_view_size + 1
        else:
            assert False, "invalid agent direction"
        botX = topX + self.agent_view_size
        botY = topY + self.agent_view_size
        return (topX, topY, botX, botY)

    def step(self, action):
        self.step_count += 1
        reward = 0
        done = False
        fwd_pos = self.front_pos
        fwd_cell = self.grid.get(*fwd_pos)
        if action == self.actions.left:
            self.agent_dir -= 1
            if self.agent_dir < 0:
                self.agent_dir += 4
        elif action == self.actions.right:
            self.agent_dir = (self.agent_dir + 1) % 4
        elif action == self.actions.forward:
            if fwd_cell is None or fwd_cell.can_overlap():
                self.agent_pos = fwd_pos
            if fwd_cell is not None and fwd_cell.type == 'goal':
                done = True
                reward = self._reward()
            if fwd_cell is not None and fwd_cell.type == 'lava':
                done = True
        elif action == self.actions.pickup:
            if fwd_cell and fwd_cell.can_pickup():
                if self.carrying is None:
                    self.carrying = fwd_cell
                    self.carrying.cur_pos = np.array([-1, -1])
                    self.grid.set(*fwd_pos, None)
        elif action == self.actions.drop:
            if not fwd_cell and self.carrying:
                self.grid.set(*fwd_pos, self.carrying)
                self.carrying.cur_pos = fwd_pos
                self.carrying = None
        elif action == self.actions.toggle:
            if fwd_cell:
                fwd_cell.toggle(self, fwd_pos)
        elif action == self.actions.done:
            pass
        else:
            assert False, "unknown action"
        if self.step_count >= self.max_steps:
            done = True
        obs = self.gen_obs()
        return obs, reward, done, {}

    def gen_obs_grid(self):
        topX, topY, botX, botY = self.get_view_exts()
        grid = self.grid.slice(topX, topY, self.agent_view_size, self.agent_view_size)
        for i in range(self.agent_dir + 1):
            grid = grid.rotate_left()
        if not self.see_through_walls:
            vis_mask = grid.process_vis(agent_pos=(self.agent_view_size // 2, self.agent_view_size - 1))
        else:
            vis_mask = np.ones(shape=(grid.width, grid.height), dtype=bool)
        agent_pos = grid.width // 2, grid.height - 1
        if self.carrying:
            grid.set(*agent_pos, self.carrying)
        else:
            grid.set(*agent_pos, None)
        return grid, vis_mask

    def gen_obs(self):
        grid, vis_mask = self.gen_obs_grid()
        image = grid.encode(vis_mask)
        assert hasattr(self, 'mission'), "environments must define a textual mission string"
        obs = {'image': image, 'direction': self.agent_dir, 'mission': self.mission}
        return obs

    def render(self, mode='human', close=False, highlight=True, tile_size=TILE_PIXELS):
        return None

    def close(self):
        return
