"""``python -m model.count_co_events`` -- same command line as the reference stage; the work is done
by otto_recommender_b200 (hand-written sm_100a CUDA behind libottocov.so)."""
import logging

from otto_recommender_b200.count_co_events import (  # noqa: F401
    concat_files_w_stats, count_co_events, count_co_events_all_files, count_population, main)

if __name__ == "__main__":
    logging.basicConfig(format="%(asctime)s - %(name)s - %(levelname)s - %(message)s", level=logging.DEBUG)
    main()
