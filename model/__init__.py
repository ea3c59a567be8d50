"""Entry-point shim package: keeps ``python -m model.count_co_events`` (reference README.md:307-311)
working on top of otto_recommender_b200.  Unlike the reference's model/__init__.py it imports nothing
heavy (the reference pulls in IPython and plotly here)."""
