from ..plugins import register_film
from .phasor_hdr_film import PhasorHDRFilm
from .transient_hdr_film import TransientHDRFilm

register_film("transient_hdr_film", lambda props: TransientHDRFilm(props))
register_film("phasor_hdr_film", lambda props: PhasorHDRFilm(props))      # phasor_hdr_film.py:283
