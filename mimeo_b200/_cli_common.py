"""Argument groups shared by the three sub-commands (names, types and defaults follow run_self.py:50-164,
run_interspecies.py:56-168 and run_map.py:53-185 of the reference byte for byte)."""
import argparse

from . import __version__


def add_version(parser):
    parser.add_argument('--version', action='version', version=f'%(prog)s {__version__}', help='Show program version and exit.')


def add_loglevel(parser):
    parser.add_argument('--loglevel', type=str, default='INFO', choices=['DEBUG', 'INFO', 'WARNING', 'ERROR', 'CRITICAL'],
                        help='Set the logging level.')

