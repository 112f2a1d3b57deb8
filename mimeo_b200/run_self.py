"""`mimeo self` (host mirror of src/mimeo/run_self.py): CLI surface unchanged, body re-pointed at the GPU engine."""
import argparse
import logging
import os
import shutil
from typing import List

from ._cli_common import add_loglevel, add_version
from .logs import init_logging
from .utils import chromlens, get_all_pairs, run_cmd, set_paths
from .wrappers import self_LZ_cmds


def mainArgs() -> argparse.Namespace:
    parser = argparse.ArgumentParser(
        description='Internal repeat finder. Mimeo-self aligns a genome to itself and extracts high-identity segments above an coverage threshold.',
        prog='mimeo-self')
    add_version(parser)
    parser.add_argument('--adir', type=str, default=None, help='Name of directory containing sequences from genome. Write split files here if providing genome as multifasta.')
    parser.add_argument('--afasta', type=str, default=None, help='Genome as multifasta.')
    parser.add_argument('-r', '--recycle', action='store_true', help='Use existing alignment "--outfile" if found.')
    parser.add_argument('-d', '--outdir', type=str, default=None, help='Write output files to this directory. (Default: cwd)')
    parser.add_argument('--gffout', type=str, default='mimeo-self_repeats.gff3', help='Name of GFF3 annotation file.')
    parser.add_argument('--outfile', type=str, default='mimeo_alignment.tab', help='Name of alignment result file.')
    parser.add_argument('--verbose', action='store_true', default=False, help='If set report alignment stage counters.')
    parser.add_argument('--label', type=str, default='Self_Repeat', help='Set annotation TYPE field in gff.')
    parser.add_argument('--prefix', type=str, default='Self_Repeat', help='ID prefix for internal repeats.')
    parser.add_argument('--keeptemp', action='store_true', default=False, help='If set do not remove temp files.')
    parser.add_argument('--lzpath', type=str, default='lastz', help='Accepted for compatibility; alignment runs on the GPU.')
    parser.add_argument('--bedtools', type=str, default='bedtools', help='Accepted for compatibility; coverage runs on the GPU.')
    parser.add_argument('--minIdt', type=int, default=60, help='Minimum alignment identity to report.')
    parser.add_argument('--minLen', type=int, default=100, help='Minimum alignment length to report.')
    parser.add_argument('--minCov', type=int, default=3, help='Minimum depth of aligned segments to report repeat feature.')
    parser.add_argument('--hspthresh', type=int, default=3000, help='Set HSP min score threshold.')
    parser.add_argument('--intraCov', type=int, default=5, help='Minimum depth of aligned segments from same scaffold to report feature. Used if "--strictSelf" mode is selected.')
    parser.add_argument('--strictSelf', action='store_true', help='If set process same-scaffold alignments separately with option to use higher "--intraCov" threshold.')
    add_loglevel(parser)
    return parser.parse_args()


def main() -> None:
    args = mainArgs()
    init_logging(loglevel=args.loglevel)
    logging.info('Starting self-alignment workflow.')
    logging.debug('Command line arguments: %s', args)
    adir_path, bdir_path, outdir, outtab, gffout, tempdir = set_paths(
        adir=args.adir, afasta=args.afasta, outdir=args.outdir, outtab=args.outfile, gffout=args.gffout, suppresBdir=True)
    pairs = get_all_pairs(Adir=adir_path, Bdir=bdir_path)
    lenPathA = os.path.join(outdir, 'A_gen_lens.txt')
    chromlens(seqDir=adir_path, outfile=lenPathA)
    cmds: List[str] = self_LZ_cmds(
        lzpath=args.lzpath, bdtlsPath=args.bedtools, pairs=pairs, Adir=adir_path, Bdir=bdir_path, outtab=outtab, outgff=gffout,
        minIdt=args.minIdt, minLen=args.minLen, hspthresh=args.hspthresh, minCov=args.minCov, intraCov=args.intraCov,
        splitSelf=args.strictSelf, AchrmLens=lenPathA, reuseTab=args.recycle, label=args.label, prefix=args.prefix)
    logging.info('Running alignments...')
    run_cmd(cmds, verbose=args.verbose, keeptemp=args.keeptemp)
    if tempdir and os.path.isdir(tempdir) and not args.keeptemp:
        shutil.rmtree(tempdir)
    logging.info('Finished!')
