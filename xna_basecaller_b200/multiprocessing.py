"""Background iterators with the reference's names (ub-bonito/bonito/multiprocessing.py:20-24, 92-122):
thread_iter runs a generator on its own thread behind a bounded queue, which is how the stages of
crf.basecall.basecall overlap (chunking, batching, the GPU call, stitching)."""
import queue
from threading import Thread

_END = object()


class ThreadIterator(Thread):
    """Drain `iterator` on a daemon thread into a queue of at most `maxsize` items."""

    def __init__(self, iterator, maxsize=1):
        super().__init__(daemon=True)
        self.iterator = iterator
        self.queue = queue.Queue(maxsize)
        self.error = None

    def run(self):
        try:
            for item in self.iterator:
                self.queue.put(item)
        except BaseException as e:      # surfaced on the consumer side instead of dying silently
            self.error = e
        self.queue.put(_END)

    def __iter__(self):
        self.start()
        while True:
            item = self.queue.get()
            if item is _END:
                break
            yield item
        if self.error is not None:
            raise self.error

    def stop(self):
        self.join()


def thread_iter(iterator, maxsize=1):
    """Take an iterator and run it on another thread."""
    return iter(ThreadIterator(iterator, maxsize=maxsize))
