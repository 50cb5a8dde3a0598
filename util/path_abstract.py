"""Base class of the site-configuration hook used by ``mypath.Path``: three static getters
(dataset root, output root, pretrained-model directory), each of which a site must override."""
import abc


def _unset(what):
    raise NotImplementedError(f"mypath.Path.{what}() is not configured for this site")


class PathAbstract(abc.ABC):
    db_root_dir = staticmethod(lambda: _unset("db_root_dir"))
    save_root_dir = staticmethod(lambda: _unset("save_root_dir"))
    models_dir = staticmethod(lambda: _unset("models_dir"))
