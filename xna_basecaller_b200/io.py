"""Output side of the basecaller: FASTQ / FASTA records and the per-read summary table, with the record formats and the
Writer / NullWriter thread contract of the reference (ub-bonito/bonito/io.py: write_fasta :69-73, write_fastq :76-84,
summary_file :148-155, summary_field_names / summary_row :158-236, CSVLogger :322-356, NullWriter :359-376, Writer
:379-445), so `bonito basecaller`-style drivers and eval scripts (eval_model.sh:102-106) read the same files.

Unaligned output only (mode 'wfq' = FASTQ, the reference's default without --reference): the SAM/BAM/CRAM modes of the
reference go through pysam + mappy, which are outside the hot path and are rejected loudly.
"""
import csv
import os
import sys
from logging import getLogger
from os.path import realpath, splitext
from threading import Thread

import numpy as np

logger = getLogger('xna_basecaller_b200')


def mean_qscore_from_qstring(qstring):
    """Mean Phred quality of a FASTQ quality string, averaged in error-probability space and capped at Q40
    (bonito/util.py:124-131).  The UB models emit a constant 'O' qstring, i.e. 40.0."""
    if len(qstring) == 0:
        return 0.0
    q = np.frombuffer(qstring.encode('ascii'), dtype=np.uint8).astype(np.float64) - 33.0
    err = np.power(10.0, -q / 10.0).mean()
    return float(-10.0 * np.log10(max(err, 1e-4)))


def write_fasta(header, sequence, fd=sys.stdout):
    fd.write('>%s\n%s\n' % (header, sequence))


def write_fastq(header, sequence, qstring, fd=sys.stdout, tags=None, sep='\t'):
    """One four-line FASTQ record; tags (SAM-style strings) follow the read id on the header line."""
    head = '@%s' % header if tags is None else '@%s %s' % (header, sep.join(tags))
    fd.write('%s\n%s\n+\n%s\n' % (head, sequence, qstring))


def summary_file():
    """summary.tsv next to a redirected stdout, else in the working directory."""
    try:
        stdout = realpath('/dev/fd/1')
    except OSError:
        stdout = ''
    if sys.stdout.isatty() or stdout.startswith('/proc') or not stdout:
        return 'summary.tsv'
    return '%s_summary.tsv' % splitext(stdout)[0]


summary_field_names = [
    'filename', 'read_id', 'run_id', 'channel', 'mux', 'start_time', 'duration', 'template_start', 'template_duration',
    'sequence_length_template', 'mean_qscore_template',
    # alignment columns: written with the reference's "no alignment" fillers when alignment is None
    'alignment_genome', 'alignment_genome_start', 'alignment_genome_end', 'alignment_strand_start',
    'alignment_strand_end', 'alignment_direction', 'alignment_length', 'alignment_num_aligned', 'alignment_num_correct',
    'alignment_num_insertions', 'alignment_num_deletions', 'alignment_num_substitutions', 'alignment_mapq',
    'alignment_strand_coverage', 'alignment_identity', 'alignment_accuracy',
]

_UNALIGNED = ['*', -1, -1, -1, -1, '*', 0, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0]


def summary_row(read, seqlen, qscore, alignment=False):
    """Row of the summary table for one read.  alignment=False: the 11 read columns only; alignment=None: plus the
    unaligned fillers.  (An actual mappy alignment is out of scope here.)"""
    attr = lambda name: getattr(read, name, '-')
    fields = [attr('filename'), read.read_id, attr('run_id'), attr('channel'), attr('mux'), attr('start'), attr('duration'),
              attr('template_start'), attr('template_duration'), seqlen, qscore]
    if alignment is None:
        fields = fields + _UNALIGNED
    elif alignment:
        raise NotImplementedError('aligned summaries need mappy; out of scope of the B200 hot path')
    return dict(zip(summary_field_names, fields))


class CSVLogger:
    """Delimited table opened for appending.  A file that already has content keeps its header row (new rows are aligned to
    it, missing keys become '-'); a new file gets the keys of the first appended row as header.  Rows are flushed in
    batches of ~100 and on close.  Behaviour of bonito/io.py:322-356."""

    FLUSH_EVERY = 100

    def __init__(self, filename, sep=','):
        self.filename, self.sep = str(filename), sep
        self.columns = self._existing_header()
        self.fh = open(self.filename, 'a', newline='')
        self.csvwriter = csv.writer(self.fh, delimiter=sep)
        self._unflushed = 0

    def _existing_header(self):
        if not os.path.exists(self.filename):
            return None
        with open(self.filename, newline='') as f:
            first = next(csv.reader(f, delimiter=self.sep), None)
        return first or None

    def set_columns(self, columns):
        if self.columns:
            raise Exception('Columns already set')
        self.columns = list(columns)
        self.csvwriter.writerow(self.columns)

    def append(self, row):
        if self.columns is None:
            self.set_columns(row.keys())
        self.csvwriter.writerow([row.get(name, '-') for name in self.columns])
        self._unflushed += 1
        if self._unflushed > self.FLUSH_EVERY:
            self.fh.flush()
            self._unflushed = 0

    def close(self):
        self.fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class NullWriter(Thread):
    """Consumes the basecall iterator and only keeps the (read_id, samples) log."""

    def __init__(self, mode, iterator, duplex=False, **kwargs):
        super().__init__()
        if duplex:
            raise NotImplementedError('duplex output is out of scope')
        self.log = []
        self.iterator = iterator

    def run(self):
        for read, res in self.iterator:
            self.log.append((read.read_id, len(read.signal)))


class Writer(Thread):
    """Thread that drains `iterator` (the (read, result) pairs of crf.basecall) into FASTQ records on `fd` and rows of the
    summary table; `log` collects (read_id, samples) for the throughput line of the CLI (cli/basecaller.py:157-161)."""

    def __init__(self, mode, iterator, aligner=None, fd=sys.stdout, duplex=False, ref_fn=None, groups=None,
                 group_key=None, summary=None):
        super().__init__()
        if mode != 'wfq' or aligner is not None or duplex:
            raise NotImplementedError("only unaligned FASTQ output (mode 'wfq') is provided; SAM/BAM/CRAM, alignment and "
                                      'duplex go through pysam/mappy in the reference and are out of scope here')
        self.fd, self.mode, self.iterator, self.group_key = fd, mode, iterator, group_key
        self.summary = summary
        self.log = []

    def run(self):
        with CSVLogger(self.summary or summary_file(), sep='\t') as summary:
            for read, res in self.iterator:
                seq = res['sequence']
                qstring = res.get('qstring', '*')
                mean_qscore = res.get('mean_qscore', mean_qscore_from_qstring(qstring))
                tags = ['RG:Z:%s_%s' % (getattr(read, 'run_id', '-'), self.group_key), 'qs:i:%d' % round(mean_qscore)]
                if hasattr(read, 'tagdata'):
                    tags.extend(read.tagdata())
                tags.extend(res.get('mods', []))
                if len(seq):
                    write_fastq(read.read_id, seq, qstring, fd=self.fd, tags=tags)
                    summary.append(summary_row(read, len(seq), mean_qscore, alignment=False))
                    self.log.append((read.read_id, len(read.signal)))
                else:
                    logger.warning('> skipping empty sequence %s', read.read_id)
