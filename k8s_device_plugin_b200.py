"""Import shim: the package directory is named after the reference repo (``k8s-device-plugin_b200``), which is not
a valid Python identifier. Importing ``k8s_device_plugin_b200`` executes that directory's ``__init__.py`` as this
module and makes its sub-modules importable (``k8s_device_plugin_b200.plugin`` ...)."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_os.path.join(_here, "k8s-device-plugin_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"))
