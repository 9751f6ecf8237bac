from ..plugins import register_integrator
from .common import TransientADIntegrator
from .transientpath import TransientPath

register_integrator("transient_path", lambda props: TransientPath(props))
