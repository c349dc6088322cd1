"""Constants of the env contract (reference config.py:1-11); values are part of the drop-in API."""
import os

PATH = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))

DEFAULT_DEVICE = 'cuda'

FOOD_CHANNEL = 0
HEAD_CHANNEL = 1
BODY_CHANNEL = 2

EPS = 1e-6
