"""Feature-set switches of the codec -- same names and defaults as the reference's constants.py:3-14.

They are module globals, not part of the bitstream: the decoder must run with the values the encoder used.
"""
USE_COORDINATES = False   # prepend (y, x) in [-1, 1] to every pixel's feature vector
EMBEDDING = False         # expand each coordinate with N_FREQ sin/cos pairs of frequency SIGMA**k * pi
SIGMA = 1.4
N_FREQ = 12
USE_COLORS = True         # (2D+1)^2 neighbourhood of normalised MSB values per band
RELATIVE = True           # subtract the centre value from the neighbourhood (only when D > 0)
