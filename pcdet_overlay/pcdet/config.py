"""Config plumbing of the reference driver (pcdet/config.py:16-82): yaml -> EasyDict with `_BASE_CONFIG_` merge and
`--set KEY VALUE` overrides.  Same function names and behaviour."""
from ast import literal_eval
from pathlib import Path

import yaml

from pcseqlearning_b200.utils import EasyDict


def cfg_from_list(cfg_list, config):
    """Set config keys via a flat [key, value, key, value, ...] list (command line --set)."""
    assert len(cfg_list) % 2 == 0
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        key_list = k.split(".")
        d = config
        for subkey in key_list[:-1]:
            assert subkey in d, "NotFoundKey: %s" % subkey
            d = d[subkey]
        subkey = key_list[-1]
        try:
            value = literal_eval(v)
        except Exception:
            value = v
        if subkey in d and type(value) != type(d[subkey]) and isinstance(d[subkey], EasyDict):
            for src in value.split(","):
                cur_key, cur_val = src.split(":")
                d[subkey][cur_key] = type(d[subkey][cur_key])(cur_val)
        elif subkey in d and type(value) != type(d[subkey]) and isinstance(d[subkey], list):
            d[subkey] = [type(d[subkey][0])(x) for x in value.split(",")]
        elif subkey not in d:
            d[subkey] = value
        else:
            assert type(value) == type(d[subkey]), f"type {type(value)} does not match original type {type(d[subkey])}"
            d[subkey] = value


def merge_new_config(config, new_config):
    if "_BASE_CONFIG_" in new_config:
        with open(new_config["_BASE_CONFIG_"], "r") as f:
            config.update(EasyDict(yaml.safe_load(f)))
    for key, val in new_config.items():
        if not isinstance(val, dict):
            config[key] = val
            continue
        if key not in config:
            config[key] = EasyDict()
        merge_new_config(config[key], val)
    return config


def cfg_from_yaml_file(cfg_file, config):
    with open(cfg_file, "r") as f:
        merge_new_config(config=config, new_config=yaml.safe_load(f))
    return config


cfg = EasyDict()
cfg.ROOT_DIR = (Path(__file__).resolve().parent / "../").resolve()
cfg.LOCAL_RANK = 0
cfg.DATA_CONFIG = EasyDict()
