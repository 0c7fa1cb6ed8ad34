from .yaml_loader import load_telescope, build_telescope
from .config_builder import TelescopeConfigBuilder
from .scene_pack import pack_config, unpack_config, load_packed_config, packaged_config

__all__ = ["load_telescope", "build_telescope", "TelescopeConfigBuilder", "pack_config", "unpack_config", "load_packed_config", "packaged_config"]
