"""Using the B200 Boltzmann solver from an unmodified WallGo script (needs WallGo itself installed).

    import WallGo
    import wallgo_b200.dropin as b200

    b200.install(vectoriseEOM=True)     # BoltzmannSolver on the GPU + lock-step EOM.findPlasmaProfile
    manager = WallGo.WallGoManager()
    ...                                  # registerModel, setupThermodynamicsHydrodynamics, collision data
    results = manager.solveWall(WallGo.WallSolverSettings(bIncludeOffEquilibrium=True))

The script below runs WallGo's own Yukawa example that way when WallGo and its Models/ directory are available
(pass the path of Models/Yukawa/yukawa.py); collision files are read by WallGo's own loader.
"""
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        return
    import wallgo_b200.dropin as b200

    b200.install(vectoriseEOM=True)
    runpy.run_path(sys.argv[1], run_name="__main__")


if __name__ == "__main__":
    main()
