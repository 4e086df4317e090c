"""TEST INFRASTRUCTURE — maps the reference's secondary-side object graph onto PlantState /
PlantParams field names (filled in subsystem by subsystem as the restatement grows)."""


def extract(sim, d):
    if not (sim.enable_secondary and sim.secondary_physics is not None):
        return


def extract_params(sim, d):
    return
