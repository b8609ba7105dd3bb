"""Host utilities with the reference's names (scikit_tt/utils.py): progress bar and timer."""
import sys
import time

_UP2 = '\x1b[2A\r'
_UL, _OFF, _GREEN, _GREY = '\x1b[4m', '\x1b[0m', '\x1b[42m', '\x1b[100m'


def progress(text, percent, cpu_time=0, show=True, width=47):
    """Two-line ANSI progress display (utils.py:37-94); returns time.time() when percent == 0."""
    if show:
        if percent > 0:
            sys.stdout.write(_UP2)
        bar = width - 6
        filled = int(bar * (int(percent) / 100))
        label = '100%' if percent == 100 else '%.1f%%' % percent
        sys.stdout.write(_UL + text.ljust(width) + _OFF + '\n')
        sys.stdout.write(_GREEN + _UL + ' ' * filled + _OFF + _GREY + _UL + ' ' * (bar - filled) + _OFF)
        sys.stdout.write(_UL + label.rjust(6) + _OFF + '\n')
        sys.stdout.write('CPU time: %.1fs ' % cpu_time)
        sys.stdout.flush()
        if percent == 100:
            sys.stdout.write('\n\n')
    if percent == 0:
        return time.time()


class timer(object):
    """`with timer() as t: ...; t.elapsed` (utils.py:97-109)."""

    def __enter__(self):
        self.start_time = time.time()
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.elapsed = time.time() - self.start_time
