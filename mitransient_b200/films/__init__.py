from ..plugins import register_film
from .transient_hdr_film import TransientHDRFilm

register_film("transient_hdr_film", lambda props: TransientHDRFilm(props))
