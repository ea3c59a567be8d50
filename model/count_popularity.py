"""``python -m model.count_popularity`` -- same command line as the reference stage; the work is done by
otto_recommender_b200 (hand-written sm_100a CUDA behind libottocov.so)."""
import logging

from otto_recommender_b200.count_popularity import count_popularity, join_clusters, main, run  # noqa: F401

if __name__ == "__main__":
    logging.basicConfig(format="%(asctime)s - %(name)s - %(levelname)s - %(message)s", level=logging.DEBUG)
    main()
