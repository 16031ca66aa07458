"""Stand-in for the absent `ruamel.yaml` on top of PyYAML."""
import yaml as _yaml


class YAML:
    def __init__(self, typ=None, pure=False):
        self.default_flow_style = False

    def load(self, stream):
        return _yaml.safe_load(stream)

    def dump(self, data, stream=None):
        return _yaml.safe_dump(data, stream)
