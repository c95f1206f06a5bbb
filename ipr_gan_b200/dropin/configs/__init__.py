"""``configs.Config``: attribute + item access over nested YAML dictionaries, with the reference's
surface (configs/__init__.py:4-43: parse, [], get, to_dict, to_yaml, str as sorted JSON).
The reference's YAML files load unchanged through ``Config.parse``; ``configs.presets`` builds the
protected-DCGAN schema in code for synthetic runs (no dataset or watermark files needed)."""
import json

import yaml


class Config(object):
    def __init__(self, entries):
        for key, value in entries.items():
            self.__dict__[key] = Config(value) if type(value) is dict else value

    @classmethod
    def parse(cls, fpath):
        with open(fpath, "r") as fh:
            return cls(yaml.safe_load(fh))

    def __getitem__(self, key):
        return self.__dict__[key]

    def __setitem__(self, key, value):
        self.__dict__[key] = value

    def get(self, key, default=None):
        return self.__dict__.get(key, default)

    def _plain(self):
        return {k: (v._plain() if isinstance(v, Config) else v) for k, v in self.__dict__.items()}

    def __str__(self):
        return json.dumps(self._plain(), indent=2, sort_keys=True)

    def to_dict(self):
        return json.loads(str(self))

    def to_yaml(self):
        return yaml.safe_dump(self.to_dict())
