from .demethify import main

main()
