"""Empty stand-in package so that `from tensorflow.examples.tutorials.mnist import input_data` (utils/func_utils.py:28) imports; no dataset exists here."""
