from ..plugins import register_sensor
from .nloscapturemeter import NLOSCaptureMeter, NLOSSensor

register_sensor("nlos_capture_meter", lambda props: NLOSCaptureMeter(props))
