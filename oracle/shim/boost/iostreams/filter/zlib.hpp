// empty stand-in: the reference's subcommand headers include this and use nothing from it
