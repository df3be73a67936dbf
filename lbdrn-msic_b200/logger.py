"""Root logger to stdout + `<folder>/<file>` with the reference's line format (reference logger.py:9-18);
`results_summary.py` of the reference scrapes these lines, so the format is part of the interface."""
import logging
import os
import sys

log = None


def create_logger(exp_folder, file_name, log_file_only=False):
    global log
    handlers = []
    if not log_file_only:
        handlers.append(logging.StreamHandler(sys.stdout))
    if file_name:
        path = os.path.join(exp_folder, file_name)
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        handlers.append(logging.FileHandler(path, mode="w"))
    for h in list(logging.root.handlers):
        logging.root.removeHandler(h)
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s] %(message)s", handlers=handlers)
    log = logging.getLogger()


def destroy_logger():
    for h in list(log.handlers):
        h.close()
        log.removeHandler(h)
