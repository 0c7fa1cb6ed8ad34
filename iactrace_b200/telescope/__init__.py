from .telescope import Telescope
from .mirrors import Mirror, MirrorGroup, AsphericDiskMirrorGroup, AsphericPolygonMirrorGroup, group_mirrors

__all__ = ["Telescope", "Mirror", "MirrorGroup", "AsphericDiskMirrorGroup", "AsphericPolygonMirrorGroup", "group_mirrors"]
