"""Config parsing and angle helpers with the reference's names (envs/utils/utils.py:12-27,144-154)."""
import os

import torch
import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "configs")


def parse_config(filename):
    """yaml -> class with the keys as attributes, read later with getattr(config, key, default).
    `filename` is a bundled config name ('heading', 'control', 'tracking') or a path to a yaml file
    (e.g. the reference's own envs/configs/heading.yaml)."""
    path = filename if os.path.isfile(str(filename)) else os.path.join(CONFIG_DIR, f"{filename}.yaml")
    assert os.path.exists(path), f"config path {path} does not exist"
    with open(path, "r", encoding="utf-8") as f:
        data = yaml.load(f, Loader=yaml.FullLoader)
    return type("EnvConfig", (object,), data)


def _t2n(x):
    return x.detach().cpu().numpy()


def wrap_2PI(angle):
    res = angle % (2 * torch.pi)
    return res + 2 * torch.pi * (res < 0)


def wrap_PI(angle):
    res = wrap_2PI(angle)
    return res - 2 * torch.pi * (res > torch.pi)
